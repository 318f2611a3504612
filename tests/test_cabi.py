"""CPU tests of the drop-in boundary: libscmgan.so loads without a GPU/driver, exports every symbol that
include/scmgan.h declares, the ctypes table covers them all, and argument validation reports errors through the
C ABI's return codes (no compute is launched here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scmgan.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from scm_gan_b200 import _lib
    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scmgan_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("scmgan_conv3x3_fwd", "scmgan_conv3x3_dgrad", "scmgan_conv3x3_wgrad", "scmgan_spectral_norm_fwd",
                 "scmgan_spectral_norm_bwd", "scmgan_clip_adam", "scmgan_bce_logits", "scmgan_reward_head_fwd",
                 "scmgan_reward_head_bwd", "scmgan_version", "scmgan_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    handle = C.CDLL(lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in scmgan.h but not exported"


def test_ctypes_table_matches_header(lib):
    assert sorted(lib.SIGNATURES) == declared_functions()


def test_version_and_error_reporting_without_gpu(lib):
    l = lib.lib()
    assert l.scmgan_version() >= 100
    assert l.scmgan_conv3x3_fwd(None, None) == -1  # SCMGAN_EINVAL
    assert b"null descriptor" in l.scmgan_last_error()
    d = lib.ConvDesc()
    d.B, d.H, d.W, d.cin, d.n = 1, 4, 4, 20, 16  # cin not a multiple of 16
    d.x, d.w = 16, 16
    assert l.scmgan_conv3x3_fwd(C.byref(d), None) == -1
    assert b"multiple of 16" in l.scmgan_last_error()
    w = lib.WgradDesc()
    assert l.scmgan_conv3x3_wgrad(C.byref(w), None) == -1
    assert l.scmgan_clip_adam(1, None, 1e-4, 0.9, 0.999, 1e-8, 1, None, 1.0, None) == -1
    assert l.scmgan_spectral_norm_fwd(0, None, None) == -1


def test_struct_layouts_match_header_sizes(lib, tmp_path):
    """Every ctypes.Structure of the binding has the size (and the offset of its last field) the C compiler gives
    the corresponding struct of include/scmgan.h."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"PackJob": "scmgan_pack_job", "ConvDesc": "scmgan_conv_desc", "WgradReduceJob": "scmgan_wgrad_reduce_job",
             "WgradDesc": "scmgan_wgrad_desc", "SnLayer": "scmgan_sn_layer", "SnBwdLayer": "scmgan_sn_bwd_layer",
             "CsrnSweepDesc": "scmgan_csrn_sweep_desc", "AdamChunk": "scmgan_adam_chunk",
             "ReplayDesc": "scmgan_replay_desc", "DecoderBceDesc": "scmgan_decoder_bce_desc"}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    body = "".join(f'  printf("{py} %zu %zu\\n", sizeof({c}), offsetof({c}, {getattr(lib, py)._fields_[-1][0]}));\n'
                   for py, c in pairs.items())
    # ... and every single field offset (a reordered or mistyped field in the middle keeps size and last offset)
    body += "".join(f'  printf("{py}.{f[0]} %zu 0\\n", offsetof({c}, {f[0]}));\n'
                    for py, c in pairs.items() for f in getattr(lib, py)._fields_)
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "scmgan.h"\nint main(void) {\n' + body +
                   "  return 0; }\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = {out[i]: (int(out[i + 1]), int(out[i + 2])) for i in range(0, len(out), 3)}
    for py in pairs:
        st = getattr(lib, py)
        last = getattr(st, st._fields_[-1][0])
        assert got[py] == (C.sizeof(st), last.offset), (py, got[py], C.sizeof(st), last.offset)
        for f in st._fields_:
            assert got[f"{py}.{f[0]}"][0] == getattr(st, f[0]).offset, (py, f[0])


def test_header_is_plain_c(tmp_path):
    """include/scmgan.h is the drop-in boundary: it must compile as C99 (and C++) with no torch / CUDA headers."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "scmgan.h"\n'
                   "int main(void) { scmgan_conv_desc c; scmgan_wgrad_desc d; scmgan_wgrad_reduce_job j;\n"
                   "  (void)c; (void)d; (void)j; return scmgan_version() > 0 ? 0 : 1; }\n")
    for cc, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++17", "c++")):
        if shutil.which(cc) is None:
            pytest.skip(f"{cc} not available")
        subprocess.run([cc, std, "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-fsyntax-only", "-x", lang,
                        str(src)], check=True)
