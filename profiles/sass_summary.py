"""Per-kernel SASS evidence of the built library: how many tcgen05 MMAs (UTCHMMA/UTCQMMA...),
TMA loads (UTMALDG), TMA stores (UTMASTG), TMEM loads/stores (LDTM/STTM), bulk copies (UBLKCP) and cp.async (LDGSTS)
each kernel contains, plus its instruction count.  Runs without a GPU.

    python profiles/sass_summary.py [scm_gan_b200/lib/libscmgan.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scm_gan_b200", "lib", "libscmgan.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UBLKCP", "LDGSTS",
             "SYNCS", "UCGABAR", "ACQBULK"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cur = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_insts"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[cur][mn] += 1
demangled = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions per kernel")
print("# UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, "
      "UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ops, ACQBULK = griddepcontrol.wait (PDL)")
hdr = ["insts"] + [m for m in MNEMONICS]
print("kernel," + ",".join(hdr))
tot = collections.Counter()
for (mangled, c), name in zip(counts.items(), demangled):
    name = re.sub(r"^void ", "", name)
    depth, cut = 0, len(name)   # drop the trailing (parameter list): scan back to its opening parenthesis
    for i in range(len(name) - 1, -1, -1):
        depth += (name[i] == ")") - (name[i] == "(")
        if depth == 0:
            cut = i
            break
    name = name[:cut].replace("(int)", "")
    print(f"\"{name}\"," + ",".join(str(c[k]) for k in ["_insts"] + MNEMONICS))
    tot.update(c)
print("TOTAL," + ",".join(str(tot[k]) for k in ["_insts"] + MNEMONICS))
