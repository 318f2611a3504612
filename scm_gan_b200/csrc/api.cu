// C-ABI entry points of libscmgan.so (see include/scmgan.h).  Host-side only: argument validation,
// TMA tensor-map encoding, launch-geometry selection, kernel launches.  No allocation, no synchronisation.
#include "../../include/scmgan.h"
#include "conv_igemm.cuh"
#include "conv_igemm_v2.cuh"
#include "conv_igemm_v3.cuh"
#include "conv_expand.cuh"
#include "conv_wgrad.cuh"
#include "conv_wgrad_v2.cuh"
#include "conv_wgrad_narrow.cuh"
#include "csrn_sweep.cuh"
#include "elementwise.cuh"
#include "host_util.cuh"

#include <algorithm>
#include <utility>
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace scm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};  // kernels launched through this library (reported by bench.py)

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", int(e), cudaGetErrorString(e), what);
    return SCM_ECUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
        return SCM_ECUDA;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
    else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
    else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bx,
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u swz %d)", int(r), rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], swizzle_bytes);
        return SCM_ECUDA;
    }
    return SCM_OK;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: opt every kernel in once per device ordinal
// (a process that drives several GPUs through this C ABI launches on each of them).
template <class Kernel>
static cudaError_t opt_in_smem(Kernel kernel, std::atomic<unsigned long long>& done, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
#define SCM_OPT_IN_SMEM(kernel, bytes)                                   \
    do {                                                                 \
        static std::atomic<unsigned long long> _done{0};                 \
        SCM_CUDA(opt_in_smem(kernel, _done, bytes));                     \
    } while (0)


// Every kernel of this library is launched with programmatic dependent launch (PDL) enabled and starts with
// griddepcontrol.wait (pdl_sync() in ptx.cuh) before its first access to global memory: the next kernel of the
// stream is scheduled while the previous one drains, its prologue (barrier init, TMEM allocation, tensor-map
// prefetch) and the launch latency overlap the predecessor's tail, and it proceeds the moment the predecessor's
// memory operations are visible.  Inside a captured CUDA graph the launches become programmatic dependency edges.
// SCMGAN_NO_PDL=1 switches the attribute off (plain stream order).
static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SCMGAN_NO_PDL"); return !(e && atoi(e)); }();
    return on;
}
// Launch bookkeeping for the pre-wait weight load of conv_igemm_v3.cuh: launches of this library on `st` since the last
// scmgan_pack_weights on it.  The load is hoisted above griddepcontrol.wait only when at least one library kernel (every
// one of them starts with griddepcontrol.wait) sits between the pack and the convolution on the same stream.
// SCMGAN_NO_PREWAIT=1 switches it off.
static thread_local cudaStream_t t_pack_stream[4] = {nullptr, nullptr, nullptr, nullptr};
static thread_local int t_pack_since[4] = {1 << 20, 1 << 20, 1 << 20, 1 << 20};
static thread_local bool t_pack_used[4] = {false, false, false, false};
static void note_launch(cudaStream_t st, bool is_pack) {
    int slot = -1;
    for (int i = 0; i < 4; ++i)
        if (t_pack_used[i] && t_pack_stream[i] == st) slot = i;
    if (slot < 0) {
        if (!is_pack) return;
        for (int i = 0; i < 4 && slot < 0; ++i)
            if (!t_pack_used[i]) slot = i;
        if (slot < 0) slot = 0;  // recycle: the evicted stream becomes "unknown", for which nothing is hoisted
        t_pack_used[slot] = true; t_pack_stream[slot] = st;
    }
    t_pack_since[slot] = is_pack ? 0 : t_pack_since[slot] + 1;
}
static bool prewait_enabled() {
    static const bool off = [] { const char* e = getenv("SCMGAN_NO_PREWAIT"); return e && atoi(e); }();
    return !off;
}
static bool prewait_weights_ok(cudaStream_t st) {
    if (!prewait_enabled()) return false;
    for (int i = 0; i < 4; ++i)
        if (t_pack_used[i] && t_pack_stream[i] == st) return t_pack_since[i] >= 1;
    return false;  // unknown stream: be conservative
}

template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);  // errors are picked up by cudaGetLastError()
    note_launch(st, false);
}

constexpr int kSmemBudget = 200 * 1024;  // tiles; barriers/alignment slack on top (<= 227 KB per CTA)

template <int CK>
static int launch_igemm(const CUtensorMap& ta, const CUtensorMap& tb, const IgemmParams& P, cudaStream_t st) {
    const int stage_bytes = IgemmCfg<CK>::kATileBytes + igemm_b_tile_bytes(P.n, CK);
    int stages = std::min(kMaxStages, kSmemBudget / stage_bytes);
    if (stages < 2) {
        set_error("conv3x3: tile does not fit shared memory (n=%d ck=%d)", P.n, CK);
        return SCM_EUNSUPPORTED;
    }
    const int smem = stages * stage_bytes + 1024 + 256;
    SCM_OPT_IN_SMEM(conv3x3_igemm_kernel<CK>, 227 * 1024);
    const int grid = std::min(P.num_tiles, num_sms());
    launch_k(conv3x3_igemm_kernel<CK>, dim3(grid), dim3(kIgemmThreads), size_t(smem), st, ta, tb, P, stages);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

constexpr int kSmemMax = 227 * 1024;

// Weight-stationary kernel (conv_igemm_v2.cuh).  Returns 1 when the shape does not fit (caller falls back to v1).
template <int CK, int TPG>
static int launch_v2_inst(const CUtensorMap& ta, const CUtensorMap& tb, const IgemmParams& P, const IgemmV2Geom& G,
                          int gx, int nsplit, int smem, cudaStream_t st) {
    // one instantiation per epilogue family (conv_igemm_v2.cuh: HEAD)
    if (P.bce_ws) {
        SCM_OPT_IN_SMEM((conv3x3_igemm_v2_kernel<CK, TPG, kHeadBce>), kSmemMax);
        launch_k(conv3x3_igemm_v2_kernel<CK, TPG, kHeadBce>, dim3(dim3(gx, nsplit)), dim3(v2_threads<CK>()), size_t(smem), st, ta, tb, P, G);
    } else if (P.out_f32) {
        SCM_OPT_IN_SMEM((conv3x3_igemm_v2_kernel<CK, TPG, kHeadF32>), kSmemMax);
        launch_k(conv3x3_igemm_v2_kernel<CK, TPG, kHeadF32>, dim3(dim3(gx, nsplit)), dim3(v2_threads<CK>()), size_t(smem), st, ta, tb, P, G);
    } else {
        SCM_OPT_IN_SMEM((conv3x3_igemm_v2_kernel<CK, TPG, kHeadPlane>), kSmemMax);
        launch_k(conv3x3_igemm_v2_kernel<CK, TPG, kHeadPlane>, dim3(dim3(gx, nsplit)), dim3(v2_threads<CK>()), size_t(smem), st, ta, tb, P, G);
    }
    SCM_CUDA(cudaGetLastError());
    return SCM_OK;
}

// CTA-pair kernel (conv_igemm_v3.cuh): N = 128, Cin in {64, 128}.  Returns 1 when the shape does not qualify.
static int launch_igemm_v3(const scmgan_conv_desc* d, const IgemmParams& P0, long long rows, cudaStream_t st) {
    constexpr int CK = 64, RB = 128;
    static const char* off = getenv("SCMGAN_NO_PAIR");
    if (off && atoi(off)) return 1;
    if (d->n != 128 || d->cin % 64 != 0 || d->out_f32) return 1;
    const int chunks = d->cin / CK;
    const int n_half = d->n / 2;
    const bool streamed = 9 * chunks * n_half * RB > 150 * 1024;  // e.g. Cin = 256: weights ride in the stages
    IgemmParams P = P0;
    P.num_tiles = int((rows + 255) / 256);
    IgemmV2Geom G;
    memset(&G, 0, sizeof(G));
    G.n_total = d->n; G.n_cta = n_half; G.b_tile_bytes = n_half * RB; G.b_streamed = streamed ? 1 : 0;
    const int b_res = streamed ? 0 : ((9 * chunks * G.b_tile_bytes + 1023) & ~1023);
    const int fixed = b_res + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment slack*/;
    const int avail = kSmemMax - fixed;
    const int Wp = d->W + 2;
    int tpg = 0;
    static const char* tpg_env = getenv("SCMGAN_TPG");
    for (int cand : {9, 3}) {
        if (tpg_env && atoi(tpg_env) == 3 && cand == 9) continue;
        if (streamed && cand != 9) continue;
        const int extent = cand == 9 ? 2 * Wp + 2 : 2;
        const int R = 128 + extent;
        const int pieces = (R + 255) / 256;
        const int piece_rows = (((R + pieces - 1) / pieces) + 7) & ~7;
        const int a_part = (pieces * piece_rows * RB + 1023) & ~1023;
        const int stage_bytes = a_part + (streamed ? 9 * G.b_tile_bytes : 0);
        const int stages = std::min(6, avail / stage_bytes);
        if (stages < (cand == 9 ? 2 : 3)) continue;
        tpg = cand;
        G.a_part_bytes = a_part;
        G.loads = pieces; G.box_rows = piece_rows; G.a_stage_bytes = stage_bytes; G.num_stages = stages;
        for (int l = 0; l < pieces; ++l) { G.ld_row[l] = l * piece_rows; G.ld_smem[l] = l * piece_rows * RB; }
        for (int t = 0; t < cand; ++t) G.a_off16[t] = uint32_t(((t / 3) * Wp + (t % 3)) * RB) >> 4;
        break;
    }
    if (!tpg) return 1;
    if (P.coord_c >= 0 && !G.a_soft) {
        set_error("conv3x3: in-tile coordinate channels need the software-staged producer (cin == 16, W <= 69)");
        return SCM_EUNSUPPORTED;
    }
    G.groups = 9 / tpg;
    const int pairs = std::max(1, std::min(P.num_tiles, num_sms() / 2));
    G.tiles_stride = pairs;
    CUtensorMap ta, tb;
    {
        uint64_t dims[2] = {uint64_t(d->x_cs), uint64_t(rows)};
        uint64_t str[1] = {uint64_t(d->x_cs) * 2};
        uint32_t box[2] = {uint32_t(CK), uint32_t(G.box_rows)};
        int rc = encode_tmap_bf16(&ta, d->x, 2, dims, str, box, RB);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {uint64_t(d->cin), uint64_t(9 * d->n)};
        uint64_t str[1] = {uint64_t(d->cin) * 2};
        uint32_t box[2] = {uint32_t(CK), uint32_t(n_half)};
        int rc = encode_tmap_bf16(&tb, d->w, 2, dims, str, box, RB);
        if (rc) return rc;
    }
    const int smem = fixed + G.num_stages * G.a_stage_bytes;
    if (tpg == 9) {
        SCM_OPT_IN_SMEM((conv3x3_igemm_v3_kernel<64, 9>), kSmemMax);
        launch_k(conv3x3_igemm_v3_kernel<64, 9>, dim3(2 * pairs), dim3(kV2Threads), size_t(smem), st, ta, tb, P, G);
    } else {
        SCM_OPT_IN_SMEM((conv3x3_igemm_v3_kernel<64, 3>), kSmemMax);
        launch_k(conv3x3_igemm_v3_kernel<64, 3>, dim3(2 * pairs), dim3(kV2Threads), size_t(smem), st, ta, tb, P, G);
    }
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

template <int CK>
static int launch_igemm_v2(const scmgan_conv_desc* d, const IgemmParams& P, long long rows, cudaStream_t st) {
    constexpr int RB = CK * 2;
    const int chunks = d->cin / CK;
    const int budget_b = 150 * 1024;
    int n_cta = d->n, nsplit = 1;
    while (9 * chunks * n_cta * RB > budget_b && n_cta % 32 == 0 && n_cta > 64) { n_cta /= 2; nsplit *= 2; }
    if (9 * chunks * n_cta * RB > budget_b) return 1;
    IgemmV2Geom G;
    memset(&G, 0, sizeof(G));
    G.n_total = d->n; G.n_cta = n_cta; G.b_tile_bytes = n_cta * RB;
    const int b_res = (9 * chunks * G.b_tile_bytes + 1023) & ~1023;
    const int fixed = b_res + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment slack*/;
    const int avail = kSmemMax - fixed;
    const int Wp = d->W + 2;
    int tpg = 0;
    if (CK == 64) {
        // shifted-row reuse of one super tile: all 9 taps if two stages fit, else the 3 taps of one filter row
        static const char* tpg_env = getenv("SCMGAN_TPG");
        for (int cand : {9, 3}) {
            if (tpg_env && atoi(tpg_env) == 3 && cand == 9) continue;
            const int extent = cand == 9 ? 2 * Wp + 2 : 2;
            const int R = 128 + extent;
            const int pieces = (R + 255) / 256;
            const int piece_rows = (((R + pieces - 1) / pieces) + 7) & ~7;
            const int stage_bytes = (pieces * piece_rows * RB + 1023) & ~1023;
            const int stages = std::min(6, avail / stage_bytes);
            if (stages < 2) continue;
            tpg = cand;
            G.loads = pieces; G.box_rows = piece_rows; G.a_stage_bytes = stage_bytes; G.num_stages = stages;
            for (int l = 0; l < pieces; ++l) { G.ld_row[l] = l * piece_rows; G.ld_smem[l] = l * piece_rows * RB; }
            for (int t = 0; t < cand; ++t) G.a_off16[t] = uint32_t(((t / 3) * Wp + (t % 3)) * RB) >> 4;
            break;
        }
    } else {
        // narrow K chunk (Cin = 16): nine separate 128-row boxes, one per tap, share a stage (one barrier round
        // trip per tile instead of nine)
        const int tile_bytes = 128 * RB;
        const int stage_bytes = 9 * tile_bytes;
        const int stages = std::min(6, avail / stage_bytes);
        const int R = 128 + 2 * Wp + 2;
        static const char* soft_env = getenv("SCMGAN_NO_SOFT_A");
        if (chunks == 1 && R <= 272 && !(soft_env && atoi(soft_env))) {
            // one software-staged super tile per 128-row tile (see the producer warp in conv_igemm_v2.cuh)
            tpg = 9;
            G.a_soft = 1;
            G.loads = 0; G.box_rows = R;
            G.a_stage_bytes = (R * RB + 1023) & ~1023;
            G.num_stages = std::min(6, avail / G.a_stage_bytes);
            for (int t = 0; t < 9; ++t) G.a_off16[t] = uint32_t(((t / 3) * Wp + (t % 3)) * RB) >> 4;
        } else if (stages >= 2 && chunks <= 4) {
            tpg = 9;
            G.loads = 9; G.box_rows = 128; G.a_stage_bytes = stage_bytes; G.num_stages = stages;
            for (int t = 0; t < 9; ++t) {
                G.ld_row[t] = (t / 3) * Wp + (t % 3);
                G.ld_smem[t] = t * tile_bytes;
                G.a_off16[t] = uint32_t(t * tile_bytes) >> 4;
            }
        }
    }
    if (!tpg) return 1;
    G.groups = 9 / tpg;
    const int gx = std::max(1, std::min(P.num_tiles, num_sms() / nsplit));
    G.tiles_stride = gx;
    CUtensorMap ta, tb;
    {
        uint64_t dims[2] = {uint64_t(d->x_cs), uint64_t(rows)};
        uint64_t str[1] = {uint64_t(d->x_cs) * 2};
        uint32_t box[2] = {uint32_t(CK), uint32_t(G.a_soft ? 128 : G.box_rows)};
        int rc = encode_tmap_bf16(&ta, d->x, 2, dims, str, box, RB);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {uint64_t(d->cin), uint64_t(9 * d->n)};
        uint64_t str[1] = {uint64_t(d->cin) * 2};
        uint32_t box[2] = {uint32_t(CK), uint32_t(n_cta)};
        int rc = encode_tmap_bf16(&tb, d->w, 2, dims, str, box, RB);
        if (rc) return rc;
    }
    const int smem = fixed + G.num_stages * G.a_stage_bytes;
    if (CK == 64) {
        return tpg == 9 ? launch_v2_inst<64, 9>(ta, tb, P, G, gx, nsplit, smem, st)
                        : launch_v2_inst<64, 3>(ta, tb, P, G, gx, nsplit, smem, st);
    }
    return launch_v2_inst<16, 9>(ta, tb, P, G, gx, nsplit, smem, st);
}

static int conv_impl_inner(const scmgan_conv_desc* d, cudaStream_t st);

// 16-input-channel layers writing a 64/128-channel plane (conv_expand.cuh).  Returns 1 when the shape does not qualify.
static int launch_expand(const scmgan_conv_desc* d, cudaStream_t st) {
    static const char* off = getenv("SCMGAN_NO_EXPAND");
    if (off && atoi(off)) return 1;
    if (d->cin != 16 || (d->n != 64 && d->n != 128) || !d->out || d->out_f32 || d->add || d->coord_c1) return 1;
    if (d->W > 128 || d->W < 2 || d->H < 2) return 1;
    ExpandParams P;
    memset(&P, 0, sizeof(P));
    P.B = d->B; P.H = d->H; P.W = d->W; P.Hp = d->H + 2; P.Wp = d->W + 2;
    P.k = std::max(1, std::min(128 / d->W, d->H));
    P.tiles_per_img = (d->H + P.k - 1) / P.k;
    P.num_tiles = d->B * P.tiles_per_img;
    P.n = d->n;
    P.a = reinterpret_cast<const __nv_bfloat16*>(d->x); P.a_cs = d->x_cs; P.a_c_off = d->x_c_off;
    P.copy_rows = (P.k + 2) * d->W;
    P.copy_bytes = (P.copy_rows * 32 + 1023) & ~1023;
    P.scale = d->scale; P.bias = d->bias; P.bias_n = d->bias_n > 0 ? std::min(d->bias_n, d->n) : d->n;
    P.sample_bias = d->sample_bias; P.sample_scale = d->sample_scale;
    P.act = d->act; P.slope = d->slope;
    P.out = reinterpret_cast<__nv_bfloat16*>(d->out); P.out_cs = d->out_cs; P.out_c_off = d->out_c_off;
    P.wrap = d->wrap; P.gated = d->gate ? 1 : 0; P.gate_c_off = d->gate_c_off;
    P.a_fmt = d->x_fmt; P.b_fmt = d->w_fmt; P.out_fmt = d->out_fmt;
    {
        const char* dbg = getenv("SCMGAN_DEBUG");
        P.debug = dbg ? atoi(dbg) : 0;
    }
    if (d->act == SCMGAN_ACT_SIGMOID) return 1;
    const int halves = d->n / 64;
    const int b_bytes = (9 * d->n * 32 + 1023) & ~1023;
    const int fixed = b_bytes + halves * 2 * 16384 * (P.gated ? 2 : 1) + kExpEpiWarps * 32 * 4 + 256 + 1024;
    const int stages = std::min(4, (kSmemMax - fixed) / (3 * P.copy_bytes));
    if (stages < 2) return 1;
    P.num_a_stages = stages;
    // the MMA of a tile reads 128 rows from row 2 W of a copy at most: keep that inside the A region
    if ((2 * d->W + 128) * 32 > 3 * P.copy_bytes) return 1;
    const int Hp = P.Hp, Wp = P.Wp;
    CUtensorMap tb, tout, tgate;
    {
        uint64_t dims[2] = {16, uint64_t(9 * d->n)};
        uint64_t str[1] = {32};
        uint32_t box[2] = {16, uint32_t(d->n)};
        int rc = encode_tmap_bf16(&tb, d->w, 2, dims, str, box, 32);
        if (rc) return rc;
    }
    auto interior = [&](CUtensorMap* t, const void* base, int cs) -> int {
        const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(base) + (size_t(Wp) + 1) * cs;
        uint64_t dims[4] = {uint64_t(cs), uint64_t(d->W), uint64_t(d->H), uint64_t(d->B)};
        uint64_t str[3] = {uint64_t(cs) * 2, uint64_t(Wp) * cs * 2, uint64_t(Hp) * Wp * cs * 2};
        uint32_t box[4] = {64, uint32_t(d->W), uint32_t(P.k), 1};
        return encode_tmap_bf16(t, bp, 4, dims, str, box, 128);
    };
    int rc = interior(&tout, d->out, d->out_cs);
    if (rc) return rc;
    if (P.gated) {
        rc = interior(&tgate, d->gate, d->gate_cs);
        if (rc) return rc;
    } else {
        tgate = tout;
    }
    const int smem = fixed + stages * 3 * P.copy_bytes;
    SCM_OPT_IN_SMEM(conv3x3_expand_kernel, kSmemMax);
    const int grid = std::min(P.num_tiles, num_sms());
    launch_k(conv3x3_expand_kernel, dim3(grid), dim3(kExpThreads), size_t(smem), st, tb, tout, tgate, P);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

static thread_local const scmgan_decoder_bce_desc* t_bce = nullptr;  // set by scmgan_decoder_bce_fwd around conv_impl

static int conv_impl(const scmgan_conv_desc* d, cudaStream_t st) {
    const int rc = conv_impl_inner(d, st);
    if (rc == SCM_OK && d->sample_out && d->rng_state && !d->uniforms) {
        // one Philox counter per 4 elements
        const unsigned long long n = ((unsigned long long)d->B * d->n_valid * d->H * d->W + 3) / 4;
        launch_k(rng_advance_kernel, dim3(1), dim3(1), size_t(0), st, d->rng_state, n);
        SCM_CUDA(cudaGetLastError());
        ++g_launches;
    }
    return rc;
}

static int conv_impl_inner(const scmgan_conv_desc* d, cudaStream_t st) {
    SCM_REQUIRE(d != nullptr, "conv3x3: null descriptor");
    SCM_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "conv3x3: bad geometry B=%d H=%d W=%d", d->B, d->H, d->W);
    SCM_REQUIRE(d->x && d->w, "conv3x3: null input/weight pointer");
    SCM_REQUIRE(d->cin > 0 && d->cin % 16 == 0, "conv3x3: cin=%d must be a positive multiple of 16", d->cin);
    SCM_REQUIRE(d->n >= 16 && d->n <= 256 && d->n % 16 == 0, "conv3x3: n=%d must be a multiple of 16 in [16,256]",
                d->n);
    SCM_REQUIRE(d->x_cs % 8 == 0 && d->x_c_off % 8 == 0 && d->x_c_off + d->cin <= d->x_cs,
                "conv3x3: bad input channel window (cs=%d off=%d cin=%d)", d->x_cs, d->x_c_off, d->cin);
    SCM_REQUIRE(d->out || d->out_f32, "conv3x3: no output requested");
    if (d->out)
        SCM_REQUIRE(d->out_cs % 8 == 0 && d->out_c_off % 8 == 0 && d->out_c_off + d->n <= d->out_cs,
                    "conv3x3: bad output channel window (cs=%d off=%d n=%d)", d->out_cs, d->out_c_off, d->n);
    if (d->add)
        SCM_REQUIRE(d->add_cs % 8 == 0 && d->add_c_off % 8 == 0 && d->add_c_off + d->n <= d->add_cs,
                    "conv3x3: bad add channel window");
    if (d->gate)
        SCM_REQUIRE(d->gate_cs % 8 == 0 && d->gate_c_off % 8 == 0 && d->gate_c_off + d->n <= d->gate_cs,
                    "conv3x3: bad gate channel window");
    if (d->out_f32) SCM_REQUIRE(d->n_valid > 0 && d->n_valid <= d->n, "conv3x3: bad n_valid=%d", d->n_valid);
    SCM_REQUIRE(!d->sample_out || d->out_f32, "conv3x3: sample_out requires out_f32");

    IgemmParams P;
    memset(&P, 0, sizeof(P));
    P.B = d->B; P.H = d->H; P.W = d->W; P.Hp = d->H + 2; P.Wp = d->W + 2;
    const long long rows = (long long)d->B * P.Hp * P.Wp;
    SCM_REQUIRE(rows < (1LL << 31) - 256, "conv3x3: plane too large");
    P.rows = int(rows);
    P.num_tiles = int((rows + 127) / 128);
    P.n = d->n;
    const int CK = (d->cin % 64 == 0) ? 64 : 16;
    P.cin_chunks = d->cin / CK;
    P.a_c_off = d->x_c_off;
    P.a = reinterpret_cast<const __nv_bfloat16*>(d->x); P.a_cs = d->x_cs;
    P.bias_n = d->bias_n > 0 ? std::min(d->bias_n, d->n) : d->n;
    P.scale = d->scale; P.bias = d->bias; P.sample_bias = d->sample_bias; P.act = d->act; P.slope = d->slope;
    P.sample_scale = d->sample_scale;
    P.out = reinterpret_cast<__nv_bfloat16*>(d->out); P.out_cs = d->out_cs; P.out_c_off = d->out_c_off;
    P.wrap = d->wrap;
    P.add = reinterpret_cast<const __nv_bfloat16*>(d->add); P.add_cs = d->add_cs; P.add_c_off = d->add_c_off;
    P.gate = reinterpret_cast<const __nv_bfloat16*>(d->gate); P.gate_cs = d->gate_cs; P.gate_c_off = d->gate_c_off;
    P.out_f32 = d->out_f32; P.n_valid = d->n_valid; P.sample_out = d->sample_out; P.uniforms = d->uniforms;
    P.rng = d->rng_state;
    SCM_REQUIRE((d->x_fmt | 1) == 1 && (d->w_fmt | 1) == 1 && (d->out_fmt | 1) == 1, "conv3x3: bad format code");
    SCM_REQUIRE(d->x_fmt == d->w_fmt, "conv3x3: input plane and packed weights must share one 16-bit format "
                                      "(tcgen05.mma kind::f16 faults on mixed bf16/fp16 operands)");
    SCM_REQUIRE(!d->add || d->out_fmt == SCMGAN_FMT_BF16, "conv3x3: `add` planes are bf16 only");
    P.a_fmt = d->x_fmt; P.b_fmt = d->w_fmt; P.out_fmt = d->out_fmt;
    {
        const char* dbg = getenv("SCMGAN_DEBUG");
        P.debug = dbg ? atoi(dbg) : 0;
    }
    P.prewait_weights = (prewait_weights_ok(st) || (d->weights_stable && prewait_enabled())) ? 1 : 0;
    P.coord_c = d->coord_c1 - 1;
    if (d->coord_c1) {
        SCM_REQUIRE(d->coord_c1 > 0 && d->coord_c1 % 2 == 1 && d->coord_c1 + 1 <= d->cin,
                    "conv3x3: coordinate channels must be an even-aligned pair inside the input window");
        SCM_REQUIRE(d->cin == 16 && d->W + 2 <= 71 && !d->wrap,
                    "conv3x3: in-tile coordinate channels need cin == 16, W <= 69 and a zero-padded plane");
    }
    if (t_bce) {
        P.bce_y = t_bce->target; P.bce_ybs = t_bce->target_bstride; P.bce_yts = t_bce->target_tstride;
        P.bce_mask = t_bce->mask; P.bce_mbs = t_bce->mask_bstride; P.bce_mts = t_bce->mask_tstride;
        P.bce_B = t_bce->B; P.bce_T = t_bce->T; P.bce_ws = t_bce->workspace;
    }

    {
        static const char* v1env = getenv("SCMGAN_IGEMM_V1");
        if (!(v1env && atoi(v1env))) {
            if (CK == 64) {
                const int rc3 = launch_igemm_v3(d, P, rows, st);
                if (rc3 <= 0) return rc3;
            } else {
                const int rce = launch_expand(d, st);
                if (rce <= 0) return rce;
            }
            const int rc = CK == 64 ? launch_igemm_v2<64>(d, P, rows, st) : launch_igemm_v2<16>(d, P, rows, st);
            if (rc <= 0) return rc;  // launched (0) or failed (<0); 1 = shape does not fit -> first-generation kernel
        }
    }
    SCM_REQUIRE(!d->coord_c1, "conv3x3: in-tile coordinate channels are not available in the first-generation kernel");
    SCM_REQUIRE(!t_bce, "decoder_bce_fwd: shape not covered by the weight-stationary kernel (cin=%d)", d->cin);
    CUtensorMap ta, tb;
    {
        uint64_t dims[2] = {uint64_t(d->x_cs), uint64_t(rows)};
        uint64_t str[1] = {uint64_t(d->x_cs) * 2};
        uint32_t box[2] = {uint32_t(CK), 128};
        int rc = encode_tmap_bf16(&ta, d->x, 2, dims, str, box, CK * 2);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {uint64_t(d->cin), uint64_t(9 * d->n)};
        uint64_t str[1] = {uint64_t(d->cin) * 2};
        uint32_t box[2] = {uint32_t(CK), uint32_t(d->n)};
        int rc = encode_tmap_bf16(&tb, d->w, 2, dims, str, box, CK * 2);
        if (rc) return rc;
    }
    return CK == 64 ? launch_igemm<64>(ta, tb, P, st) : launch_igemm<16>(ta, tb, P, st);
}

// ---- split-K reductions: launched right away, or recorded for the caller (scmgan_wgrad_desc::defer_jobs) ----------
static thread_local scmgan_wgrad_reduce_job* t_defer_jobs = nullptr;
static thread_local int t_defer_cap = 0;
static thread_local int* t_defer_count = nullptr;
static thread_local long long t_ws_off = 0;  // bytes; advanced only while deferring
static thread_local int t_dy_fmt = 0, t_x_fmt = 0;  // operand formats of the wgrad call being dispatched

// Scratch for one launch: the whole workspace when reductions run immediately (stream order protects it), a fresh
// slice while deferring.  nullptr = does not fit.
static float* ws_take(float* ws, long long ws_bytes, long long need_bytes) {
    if (!ws) return nullptr;
    if (!t_defer_jobs) return need_bytes <= ws_bytes ? ws : nullptr;
    if (t_ws_off + need_bytes > ws_bytes) return nullptr;
    float* p = ws + t_ws_off / 4;
    t_ws_off += (need_bytes + 255) & ~255LL;
    return p;
}

static int launch_reduce(const scmgan_wgrad_reduce_job& j, cudaStream_t st) {
    const int total = 9 * j.n * 32;
    const int blocks = (total + 31) / 32 + (j.ws_bias && j.db ? 1 : 0);
    if (j.lanes == 32)
        launch_k(wgrad_reduce_kernel<32>, dim3(blocks), dim3(dim3(32, 32)), size_t(0), st, j.ws, j.splits, j.n, j.g, j.g_sm, j.g_sn, j.g_st, j.flip, j.m_valid, j.n_valid, j.scale, j.ws_bias, j.db);
    else
        launch_k(wgrad_reduce_kernel<8>, dim3(blocks), dim3(dim3(32, 8)), size_t(0), st, j.ws, j.splits, j.n, j.g, j.g_sm, j.g_sn, j.g_st, j.flip, j.m_valid, j.n_valid, j.scale, j.ws_bias, j.db);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

static int emit_reduce(const float* ws, int splits, int n, float* g, long long g_sm, long long g_sn, long long g_st,
                       int flip, int m_valid, int n_valid, float scale, const float* ws_bias, float* db, int lanes,
                       cudaStream_t st) {
    scmgan_wgrad_reduce_job j{ws, splits, n, g, g_sm, g_sn, g_st, flip, m_valid, n_valid, scale, ws_bias, db, lanes};
    if (!t_defer_jobs) return launch_reduce(j, st);
    SCM_REQUIRE(*t_defer_count < t_defer_cap, "wgrad: more than %d deferred reductions", t_defer_cap);
    t_defer_jobs[(*t_defer_count)++] = j;
    return SCM_OK;
}

// Second-generation wgrad (conv_wgrad_v2.cuh): P = dY (interior view, 128 channels), Q = X (padded view, n channels),
// W % 16 == 0, workspace required.  Returns 1 when the shape does not qualify.
static int wgrad_launch_v2(int B, int H, int W, const void* pp, int p_cs, int p_c_off, const void* qp, int q_cs,
                           int q_c_off, int n, int flip, float scale, float* g, long long g_sm, long long g_sn,
                           long long g_st, int m_valid, int n_valid, float* ws, long long ws_bytes, float* db,
                           cudaStream_t st) {
    static const char* off = getenv("SCMGAN_WGRAD_V1");
    if (off && atoi(off)) return 1;
    if (!ws || W > 240 || n > 128 || n % 16 != 0) return 1;
    const int Hp = H + 2, Wp = W + 2;
    // The K dimension runs over the pixels of an image row in steps of 16.  For W % 16 != 0 (MiniPacMan: 19) the row is
    // padded to Wk pixels INSIDE the TMA boxes: dY comes through the interior view, so pixels W..Wk-1 are out of bounds
    // and arrive as zeros (they contribute nothing, whatever X holds there); the kernel only ever sees Wk.
    const int Wk = (W + 15) & ~15;
    {
        static const char* v1_ragged = getenv("SCMGAN_WGRAD_V1_RAGGED");  // A/B switch: previous kernel for W % 16 != 0
        if (Wk != W && v1_ragged && atoi(v1_ragged)) return 1;
    }
    const int BH = std::max(1, std::min(128 / Wk, H));
    const int KP = Wk * BH;
    const int q_aw = (n % 64 == 0) ? 64 : (n % 32 == 0 ? 32 : 16);
    const int q_atoms = n / q_aw;
    const int wq = (std::max(Wp, Wk + 2) + 7) & ~7;
    if (wq > 256) return 1;
    const int q_atom_bytes = (wq * q_aw * 2 + 1023) & ~1023;
    const int stage_bytes = 2 * KP * 128 + BH * q_atoms * q_atom_bytes;
    const int stages = std::min(4, (kSmemMax - 2048) / stage_bytes);
    if (stages < 2) return 1;
    WgradV2Params P;
    memset(&P, 0, sizeof(P));
    P.B = B; P.W = Wk; P.BH = BH;
    P.nby = (H + BH - 1) / BH;
    P.num_kblocks = B * P.nby;
    int splits = std::max(1, std::min(P.num_kblocks / 4, std::max(1, num_sms() / 3)));
    P.kb_per_cta = (P.num_kblocks + splits - 1) / splits;
    splits = (P.num_kblocks + P.kb_per_cta - 1) / P.kb_per_cta;
    const long long main_floats = (long long)splits * 9 * n * 128;
    if (db && 3 * n + 16 > 512) return 1;
    ws = ws_take(ws, ws_bytes, (main_floats + (long long)splits * 128) * 4);
    if (!ws) return 1;
    P.n = n; P.q_aw = q_aw; P.wq = wq; P.p_c_off = p_c_off; P.q_c_off = q_c_off; P.ws = ws;
    P.ws_bias = db ? ws + main_floats : nullptr;
    P.p_fmt = t_dy_fmt; P.q_fmt = t_x_fmt;
    {
        static const char* dbg = getenv("SCMGAN_DEBUG");
        P.debug = dbg ? (atoi(dbg) & (8 | 16)) : 0;
    }
    CUtensorMap tp, tq;
    {
        const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(pp) + (size_t(Wp) + 1) * p_cs;
        uint64_t dims[4] = {uint64_t(p_cs), uint64_t(W), uint64_t(H), uint64_t(B)};
        uint64_t str[3] = {uint64_t(p_cs) * 2, uint64_t(Wp) * p_cs * 2, uint64_t(Hp) * Wp * p_cs * 2};
        uint32_t box[4] = {64, uint32_t(Wk), uint32_t(BH), 1};
        int rc = encode_tmap_bf16(&tp, bp, 4, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {uint64_t(q_cs), uint64_t(Wp), uint64_t(Hp), uint64_t(B)};
        uint64_t str[3] = {uint64_t(q_cs) * 2, uint64_t(Wp) * q_cs * 2, uint64_t(Hp) * Wp * q_cs * 2};
        uint32_t box[4] = {uint32_t(q_aw), uint32_t(wq), 1, 1};
        int rc = encode_tmap_bf16(&tq, qp, 4, dims, str, box, q_aw * 2);
        if (rc) return rc;
    }
    SCM_OPT_IN_SMEM(conv3x3_wgrad_v2_kernel, kSmemMax);
    const int smem = stages * stage_bytes + 1024 + 1024;
    launch_k(conv3x3_wgrad_v2_kernel, dim3(dim3(splits, 3)), dim3(kWgradThreads), size_t(smem), st, tp, tq, P, stages);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return emit_reduce(ws, splits, n, g, g_sm, g_sn, g_st, flip, m_valid, n_valid, scale, P.ws_bias, db,
                       n <= 32 ? 32 : 8, st);
}

// 16-channel-on-one-side layers (conv_wgrad_narrow.cuh): all 128-channel blocks of the wide side in one launch.
//   x_is_wide = true : cout = 16; M = X (padded view), N = dY (interior view at -tap); g[m=ci][n=co]
//   x_is_wide = false: cin  = 16; M = dY (interior view), N = X (padded view at +tap);  g[m=co][n=ci]; db optional
// g is indexed g[m*g_sm + n*g_sn + tap*g_st].  Returns 1 when the shape does not qualify.
static int wgrad_launch_narrow(bool x_is_wide, int B, int H, int W, const void* xp, int x_cs, int x_c_off,
                               const void* dyp, int dy_cs, int dy_c_off, int m_blocks, int flip, float scale, float* g,
                               long long g_sm, long long g_sn, long long g_st, int m_valid, int n_valid, float* ws,
                               long long ws_bytes, float* db, cudaStream_t st) {
    static const char* off = getenv("SCMGAN_WGRAD_NO_NARROW");
    if (off && atoi(off)) return 1;
    if (!ws) return 1;
    const int Hp = H + 2, Wp = W + 2;
    const int rows_k = x_is_wide ? Hp : H;            // image rows the K dimension runs over (the M operand's view)
    const int kw = ((x_is_wide ? Wp : W) + 15) & ~15;  // pixels per image-row box
    if (kw > 256) return 1;
    int BH = 0, stage_bytes = 0;
    for (int cand : {4, 2, 1}) {
        const int sb = (cand * kw * (256 + 10 * 32) + 1023) & ~1023;
        if (2 * sb + 2048 <= kSmemMax) { BH = cand; stage_bytes = sb; break; }
    }
    if (!BH) return 1;
    const int stages = std::min(4, (kSmemMax - 2048) / stage_bytes);
    WgradNarrowParams P;
    memset(&P, 0, sizeof(P));
    P.B = B; P.Hp = Hp; P.BH = BH; P.kw = kw;
    P.nrb = (rows_k + BH - 1) / BH;
    P.num_kblocks = B * P.nrb;
    int splits = std::max(1, std::min(P.num_kblocks / 2, std::max(1, num_sms() / m_blocks)));
    P.kb_per_cta = (P.num_kblocks + splits - 1) / splits;
    splits = (P.num_kblocks + P.kb_per_cta - 1) / P.kb_per_cta;
    const long long per_block = (long long)splits * 9 * kNarrowCo * 128;
    const long long main_floats = per_block * m_blocks;
    const bool with_ones = db != nullptr && !x_is_wide;
    ws = ws_take(ws, ws_bytes, (main_floats + (with_ones ? (long long)m_blocks * splits * 128 : 0)) * 4);
    if (!ws) return 1;
    P.x_c_off = x_is_wide ? x_c_off : dy_c_off;
    P.dy_c_off = x_is_wide ? dy_c_off : x_c_off;
    P.n_sign = x_is_wide ? -1 : +1;
    P.with_ones = with_ones ? 1 : 0;
    P.ws = ws;
    P.ws_bias = with_ones ? ws + main_floats : nullptr;
    P.m_fmt = x_is_wide ? t_x_fmt : t_dy_fmt;
    P.n_fmt = x_is_wide ? t_dy_fmt : t_x_fmt;
    {
        static const char* dbg = getenv("SCMGAN_DEBUG");
        P.debug = dbg ? (atoi(dbg) & 16) : 0;
    }
    // tensor maps: padded view of X, interior view of dY; the wide one is the M operand (64-channel boxes, 128B
    // swizzle), the 16-channel one the N operand (32B swizzle)
    CUtensorMap tx, tdy;
    {
        uint64_t dims[4] = {uint64_t(x_cs), uint64_t(Wp), uint64_t(Hp), uint64_t(B)};
        uint64_t str[3] = {uint64_t(x_cs) * 2, uint64_t(Wp) * x_cs * 2, uint64_t(Hp) * Wp * x_cs * 2};
        uint32_t box[4] = {uint32_t(x_is_wide ? 64 : kNarrowCo), uint32_t(kw), uint32_t(BH), 1};
        int rc = encode_tmap_bf16(&tx, xp, 4, dims, str, box, x_is_wide ? 128 : 32);
        if (rc) return rc;
    }
    {
        const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(dyp) + (size_t(Wp) + 1) * dy_cs;
        uint64_t dims[4] = {uint64_t(dy_cs), uint64_t(W), uint64_t(H), uint64_t(B)};
        uint64_t str[3] = {uint64_t(dy_cs) * 2, uint64_t(Wp) * dy_cs * 2, uint64_t(Hp) * Wp * dy_cs * 2};
        uint32_t box[4] = {uint32_t(x_is_wide ? kNarrowCo : 64), uint32_t(kw), uint32_t(BH), 1};
        int rc = encode_tmap_bf16(&tdy, bp, 4, dims, str, box, x_is_wide ? 32 : 128);
        if (rc) return rc;
    }
    SCM_OPT_IN_SMEM(conv3x3_wgrad_narrow_kernel, kSmemMax);
    const int smem = stages * stage_bytes + 1024 + 1024;
    if (x_is_wide)
        launch_k(conv3x3_wgrad_narrow_kernel, dim3(dim3(splits, m_blocks)), dim3(kWgradThreads), size_t(smem), st, tx, tdy, P, stages);
    else
        launch_k(conv3x3_wgrad_narrow_kernel, dim3(dim3(splits, m_blocks)), dim3(kWgradThreads), size_t(smem), st, tdy, tx, P, stages);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    for (int mb = 0; mb < m_blocks; ++mb) {
        const int mv = std::min(128, m_valid - mb * 128);
        if (mv <= 0) break;
        const int rc = emit_reduce(ws + mb * per_block, splits, kNarrowCo, g + mb * 128 * g_sm, g_sm, g_sn, g_st, flip,
                                   mv, n_valid, scale,
                                   with_ones ? P.ws_bias + (long long)mb * splits * 128 : nullptr,
                                   with_ones ? db + mb * 128 : nullptr, 32, st);
        if (rc) return rc;
    }
    return SCM_OK;
}

// one wgrad launch: P = 128 channels of `pp` (view per p_interior), Q = n channels of `qp`
static int wgrad_launch(int B, int H, int W, const void* pp, int p_cs, int p_c_off, bool p_interior, const void* qp,
                        int q_cs, int q_c_off, int n, int q_sign, int flip, float scale, float* g, long long g_sm,
                        long long g_sn, long long g_st, int m_valid, int n_valid, float* ws, long long ws_bytes,
                        cudaStream_t st) {
    const int Hp = H + 2, Wp = W + 2;
    const int Wd = p_interior ? W : Wp, Hd = p_interior ? H : Hp;
    // pixel box: BW covers a row of P's view, KP = BW*BH multiple of 16, ~64 pixels
    int best_bw = 0, best_bh = 0;
    double best_cost = 1e30;
    for (int bh = 1; bh <= 16; bh *= 2) {
        int bw = Wd;
        while ((bw * bh) % 16) ++bw;
        const int kp = bw * bh;
        if (bw > 256 || kp > 128) continue;
        const int nby = (Hd + bh - 1) / bh;
        double cost = double(kp) * nby / (double(Wd) * Hd);  // padded work ratio
        if (kp < 48) cost *= 1.0 + (48 - kp) / 48.0;         // tiny K blocks amortise barriers poorly
        if (cost < best_cost) { best_cost = cost; best_bw = bw; best_bh = bh; }
    }
    if (!best_bw) {
        set_error("wgrad: unsupported width %d", Wd);
        return SCM_EUNSUPPORTED;
    }
    const int BW = best_bw, BH = best_bh, KP = BW * BH;
    const int q_aw = (n % 64 == 0) ? 64 : (n % 32 == 0 ? 32 : 16);
    const int q_atoms = n / q_aw;
    const int q_atom_bytes = (KP * q_aw * 2 + 1023) & ~1023;
    int tg = 9;
    int stages = 0;
    for (int cand : {9, 3, 1}) {
        tg = cand;
        if (tg * n > 512) continue;
        const int stage_bytes = 2 * KP * 128 + tg * q_atoms * q_atom_bytes;
        stages = std::min(4, kSmemBudget / stage_bytes);
        if (stages >= 2) break;
    }
    if (stages < 2) {
        set_error("wgrad: tile does not fit shared memory (n=%d KP=%d)", n, KP);
        return SCM_EUNSUPPORTED;
    }
    const int stage_bytes = 2 * KP * 128 + tg * q_atoms * q_atom_bytes;
    const int groups = (9 + tg - 1) / tg;

    WgradParams P;
    memset(&P, 0, sizeof(P));
    P.B = B; P.BW = BW; P.BH = BH;
    P.nby = (Hd + BH - 1) / BH;
    P.num_kblocks = B * P.nby;
    int splits = std::max(1, std::min(P.num_kblocks / 8, std::max(1, num_sms() / groups)));
    P.kb_per_cta = (P.num_kblocks + splits - 1) / splits;
    splits = (P.num_kblocks + P.kb_per_cta - 1) / P.kb_per_cta;
    P.tap0_stride = tg; P.n = n; P.q_aw = q_aw; P.p_c_off = p_c_off; P.q_c_off = q_c_off; P.q_sign = q_sign;
    P.flip = flip; P.scale = scale; P.g = g; P.g_sm = g_sm; P.g_sn = g_sn; P.g_st = g_st;
    P.m_valid = m_valid; P.n_valid = n_valid;
    P.p_fmt = p_interior ? t_dy_fmt : t_x_fmt;  // the interior-view operand is always the gradient plane
    P.q_fmt = p_interior ? t_x_fmt : t_dy_fmt;
    {
        static const char* dbg = getenv("SCMGAN_DEBUG");
        P.debug = dbg ? (atoi(dbg) & 8) : 0;
    }
    const long long ws_need = (long long)splits * 9 * n * 128 * 4;
    P.ws = ws_take(ws, ws_bytes, ws_need);

    auto make_view = [&](CUtensorMap* t, const void* base, int cs, bool interior, int box_c, int swz) -> int {
        const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(base);
        if (interior) bp += (size_t(Wp) + 1) * cs;
        uint64_t dims[4] = {uint64_t(cs), uint64_t(interior ? W : Wp), uint64_t(interior ? H : Hp), uint64_t(B)};
        uint64_t str[3] = {uint64_t(cs) * 2, uint64_t(Wp) * cs * 2, uint64_t(Hp) * Wp * cs * 2};
        uint32_t box[4] = {uint32_t(box_c), uint32_t(BW), uint32_t(BH), 1};
        return encode_tmap_bf16(t, bp, 4, dims, str, box, swz);
    };
    CUtensorMap tp, tq;
    int rc = make_view(&tp, pp, p_cs, p_interior, 64, 128);
    if (rc) return rc;
    rc = make_view(&tq, qp, q_cs, !p_interior, q_aw, q_aw * 2);
    if (rc) return rc;

    SCM_OPT_IN_SMEM(conv3x3_wgrad_kernel, 227 * 1024);
    const int smem = stages * stage_bytes + 1024 + 256;
    launch_k(conv3x3_wgrad_kernel, dim3(dim3(splits, groups)), dim3(kWgradThreads), size_t(smem), st, tp, tq, P, stages);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    if (P.ws) return emit_reduce(P.ws, splits, n, g, g_sm, g_sn, g_st, flip, m_valid, n_valid, scale, nullptr, nullptr, 8, st);
    return SCM_OK;
}

}  // namespace scm

using namespace scm;

extern "C" {

int scmgan_version(void) { return 100; }
const char* scmgan_last_error(void) { return g_err; }
int scmgan_num_sms(void) { return num_sms(); }
long long scmgan_launch_count(void) { return g_launches.load(); }

int scmgan_pack_nchw(const float* src, long long src_bstride, int C, int B, int H, int W, void* dst_plane, int Cs,
                     int c_off, int c_pad, int wrap, const float* sig, int fmt, scmgan_stream_t stream) {
    SCM_REQUIRE((fmt | 1) == 1, "pack_nchw: bad format code");
    SCM_REQUIRE(src && dst_plane, "pack_nchw: null pointer");
    SCM_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "pack_nchw: bad geometry");
    SCM_REQUIRE(Cs % 8 == 0 && c_off % 8 == 0 && c_pad % 8 == 0 && c_pad >= C && c_off + c_pad <= Cs,
                "pack_nchw: bad channel window (Cs=%d off=%d pad=%d C=%d)", Cs, c_off, c_pad, C);
    const long long rows = (long long)B * (H + 2) * (W + 2);
    const int threads = 128;
    const long long blocks = (rows + threads - 1) / threads;
    launch_k(pack_nchw_to_plane_kernel, dim3((unsigned)blocks), dim3(threads), size_t(0), (cudaStream_t)stream, src, src_bstride, C, B, H, W, reinterpret_cast<__nv_bfloat16*>(dst_plane), Cs, c_off, c_pad, wrap, sig, fmt);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_pack_coords(void* dst_plane, int Cs, int c_off, int B, int H, int W, int fmt, scmgan_stream_t stream) {
    SCM_REQUIRE((fmt | 1) == 1, "pack_coords: bad format code");
    SCM_REQUIRE(dst_plane && B > 0 && H > 0 && W > 0, "pack_coords: bad arguments");
    SCM_REQUIRE(c_off % 2 == 0 && c_off + 2 <= Cs, "pack_coords: bad channel window (Cs=%d off=%d)", Cs, c_off);
    const long long total = (long long)B * H * W;
    launch_k(pack_coords_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), size_t(0), (cudaStream_t)stream, reinterpret_cast<__nv_bfloat16*>(dst_plane), Cs, c_off, B, H, W, fmt);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_pack_weights(int count, const scmgan_pack_job* jobs, scmgan_stream_t stream) {
    SCM_REQUIRE(count >= 0 && (count == 0 || jobs), "pack_weights: bad arguments");
    for (int base = 0; base < count; base += kMaxPackJobs) {
        PackJobs J;
        memset(&J, 0, sizeof(J));
        J.count = std::min(kMaxPackJobs, count - base);
        long long max_total = 0;
        for (int i = 0; i < J.count; ++i) {
            const scmgan_pack_job& s = jobs[base + i];
            SCM_REQUIRE(s.w && s.out && s.n_pad > 0 && s.k_pad > 0 && s.n_valid <= s.n_pad && s.k_valid <= s.k_pad,
                        "pack_weights: bad job %d", base + i);
            PackJob& d = J.job[i];
            d.w = s.w; d.out = reinterpret_cast<__nv_bfloat16*>(s.out); d.sigma = s.sigma;
            d.n_pad = s.n_pad; d.k_pad = s.k_pad; d.n_valid = s.n_valid; d.k_valid = s.k_valid;
            d.s_n = s.s_n; d.s_k = s.s_k; d.k_src_off = s.k_src_off; d.flip = s.flip;
            SCM_REQUIRE(s.out_ld == 0 || s.out_ld >= s.k_pad, "pack_weights: job %d: out_ld < k_pad", base + i);
            d.out_ld = s.out_ld ? s.out_ld : s.k_pad;
            SCM_REQUIRE((s.fmt | 1) == 1, "pack_weights: job %d: bad format code", base + i);
            d.fmt = s.fmt;
            max_total = std::max(max_total, 9LL * s.n_pad * s.k_pad);
        }
        const int threads = 256;
        const int bx = int(std::min<long long>((max_total + threads - 1) / threads, 296));
        launch_k(pack_weights_kernel, dim3(dim3(bx, J.count)), dim3(threads), size_t(0), (cudaStream_t)stream, J);
        note_launch((cudaStream_t)stream, true);  // operands change: no pre-wait weight load in the very next kernel
        SCM_CUDA(cudaGetLastError());
    ++g_launches;
    }
    return SCM_OK;
}

int scmgan_conv3x3_fwd(const scmgan_conv_desc* d, scmgan_stream_t stream) { return conv_impl(d, (cudaStream_t)stream); }
int scmgan_conv3x3_dgrad(const scmgan_conv_desc* d, scmgan_stream_t stream) {
    return conv_impl(d, (cudaStream_t)stream);
}

static int wgrad_dispatch(const scmgan_wgrad_desc* d, scmgan_stream_t stream);

int scmgan_conv3x3_wgrad(const scmgan_wgrad_desc* d, scmgan_stream_t stream) {
    SCM_REQUIRE(d != nullptr, "wgrad: null descriptor");
    SCM_REQUIRE((d->dy_fmt | 1) == 1 && (d->x_fmt | 1) == 1, "wgrad: bad format code");
    SCM_REQUIRE(d->dy_fmt == d->x_fmt || (d->dy_fmt == SCMGAN_FMT_BF16 && d->x_fmt == SCMGAN_FMT_F16),
                "wgrad: supported formats are dy == x, or a bf16 gradient plane with an fp16 input plane");
    t_dy_fmt = d->dy_fmt; t_x_fmt = d->x_fmt;
    if (d->defer_jobs) {
        SCM_REQUIRE(d->defer_count && d->workspace_cursor && d->defer_cap > 0 && d->workspace,
                    "wgrad: deferred reduction needs defer_count, workspace_cursor and a workspace");
        t_defer_jobs = d->defer_jobs; t_defer_cap = d->defer_cap; t_defer_count = d->defer_count;
        t_ws_off = *d->workspace_cursor;
    }
    const int rc = wgrad_dispatch(d, stream);
    if (d->defer_jobs) {
        *d->workspace_cursor = t_ws_off;
        t_defer_jobs = nullptr; t_defer_cap = 0; t_defer_count = nullptr; t_ws_off = 0;
    }
    return rc;
}

int scmgan_wgrad_reduce(int count, const scmgan_wgrad_reduce_job* jobs, scmgan_stream_t stream) {
    SCM_REQUIRE(count >= 0 && (count == 0 || jobs), "wgrad_reduce: bad arguments");
    for (int i = 0; i < count; ++i) {
        SCM_REQUIRE(jobs[i].ws && jobs[i].g && jobs[i].splits > 0 && jobs[i].n > 0, "wgrad_reduce: bad job %d", i);
        const int rc = launch_reduce(jobs[i], (cudaStream_t)stream);
        if (rc) return rc;
    }
    return SCM_OK;
}

static int wgrad_dispatch(const scmgan_wgrad_desc* d, scmgan_stream_t stream) {
    SCM_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->dy && d->x && d->g, "wgrad: bad arguments");
    SCM_REQUIRE(d->cout % 16 == 0 && d->cin % 16 == 0 && d->cout > 0 && d->cin > 0, "wgrad: channels must be x16");
    SCM_REQUIRE(d->dy_cs % 8 == 0 && d->x_cs % 8 == 0 && d->dy_c_off % 8 == 0 && d->x_c_off % 8 == 0,
                "wgrad: bad channel strides");
    SCM_REQUIRE(d->dy_c_off + d->cout <= d->dy_cs && d->x_c_off + d->cin <= d->x_cs, "wgrad: channel window");
    cudaStream_t st = (cudaStream_t)stream;
    // bias gradient (optional): folded into the first v2 launch of every 128-channel output block, otherwise a
    // separate interior column sum over the gradient plane
    auto bias_fallback = [&](int m0, int mcount) -> int {
        const int n8 = (std::min(mcount, d->co_valid - m0) + 7) & ~7;  // db holds co_valid (rounded up to 8) floats
        return scmgan_plane_colsum(d->dy, d->dy_cs, d->dy_c_off + m0, n8, d->B, d->H, d->W, nullptr, d->db + m0, stream);
    };
    if (d->cout % 128 == 0 && d->cin == kNarrowCo) {
        const int blocks = std::min(d->cout / 128, (d->co_valid + 127) / 128);
        const int rc = wgrad_launch_narrow(false, d->B, d->H, d->W, d->x, d->x_cs, d->x_c_off, d->dy, d->dy_cs,
                                           d->dy_c_off, blocks, d->flip, d->scale, d->g, d->g_s_co, d->g_s_ci,
                                           d->g_s_tap, d->co_valid, d->ci_valid, (float*)d->workspace,
                                           d->workspace_bytes, d->db, st);
        if (rc <= 0) return rc;
    }
    if (d->cout % 128 == 0) {
        for (int m0 = 0; m0 < d->cout; m0 += 128) {
            if (m0 >= d->co_valid) break;
            bool bias_done = (d->db == nullptr);
            for (int c0 = 0; c0 < d->cin; c0 += 128) {
                const int n = std::min(128, d->cin - c0);
                const int nv = std::min(n, d->ci_valid - c0);
                if (nv <= 0) break;
                {
                    const int rc2 = wgrad_launch_v2(d->B, d->H, d->W, d->dy, d->dy_cs, d->dy_c_off + m0, d->x, d->x_cs,
                                                    d->x_c_off + c0, n, d->flip, d->scale,
                                                    d->g + m0 * d->g_s_co + c0 * d->g_s_ci, d->g_s_co, d->g_s_ci,
                                                    d->g_s_tap, std::min(128, d->co_valid - m0), nv,
                                                    (float*)d->workspace, d->workspace_bytes,
                                                    bias_done ? nullptr : d->db + m0, st);
                    if (rc2 < 0) return rc2;
                    if (rc2 == 0) { bias_done = true; continue; }
                }
                int rc = wgrad_launch(d->B, d->H, d->W, d->dy, d->dy_cs, d->dy_c_off + m0, true, d->x, d->x_cs,
                                      d->x_c_off + c0, n, +1, d->flip, d->scale,
                                      d->g + m0 * d->g_s_co + c0 * d->g_s_ci, d->g_s_co, d->g_s_ci, d->g_s_tap,
                                      std::min(128, d->co_valid - m0), nv, (float*)d->workspace, d->workspace_bytes,
                                      st);
                if (rc) return rc;
            }
            if (!bias_done) {
                int rc = bias_fallback(m0, 128);
                if (rc) return rc;
            }
        }
        return SCM_OK;
    }
    if (d->cin % 128 == 0 && d->cout <= 256) {
        if (d->cout == kNarrowCo) {
            const int blocks = std::min(d->cin / 128, (d->ci_valid + 127) / 128);
            const int rc = wgrad_launch_narrow(true, d->B, d->H, d->W, d->x, d->x_cs, d->x_c_off, d->dy, d->dy_cs,
                                               d->dy_c_off, blocks, d->flip, d->scale, d->g, d->g_s_ci, d->g_s_co,
                                               d->g_s_tap, d->ci_valid, d->co_valid, (float*)d->workspace,
                                               d->workspace_bytes, nullptr, st);
            if (rc < 0) return rc;
            if (rc == 0) {
                if (d->db) return bias_fallback(0, d->cout);
                return SCM_OK;
            }
        }
        for (int c0 = 0; c0 < d->cin; c0 += 128) {
            const int mv = std::min(128, d->ci_valid - c0);
            if (mv <= 0) break;
            int rc = wgrad_launch(d->B, d->H, d->W, d->x, d->x_cs, d->x_c_off + c0, false, d->dy, d->dy_cs,
                                  d->dy_c_off, d->cout, -1, d->flip, d->scale, d->g + c0 * d->g_s_ci, d->g_s_ci,
                                  d->g_s_co, d->g_s_tap, mv, d->co_valid, (float*)d->workspace, d->workspace_bytes,
                                  st);
            if (rc) return rc;
        }
        if (d->db) {
            int rc = bias_fallback(0, d->cout);
            if (rc) return rc;
        }
        return SCM_OK;
    }
    set_error("wgrad: one of cout (%d) / cin (%d) must be a multiple of 128", d->cout, d->cin);
    return SCM_EUNSUPPORTED;
}

/* profiling aid (not part of the documented ABI): copies the in-kernel timeline of conv_expand.cuh to the host */
int scmgan_debug_expand_timeline(unsigned long long* out_host, int n) {
    SCM_CUDA(cudaDeviceSynchronize());
    SCM_CUDA(cudaMemcpyFromSymbol(out_host, g_exp_dbg, sizeof(unsigned long long) * std::min(n, 3 * 16 * 8)));
    return SCM_OK;
}

long long scmgan_wgrad_workspace_bytes(void) {
    // upper bound over all shapes: (#CTAs of one launch) x (taps per CTA group) x 128 x n floats, n <= 128
    return (long long)(num_sms() + 8) * 9 * 128 * 128 * 4 / 3 + (2 << 20);
}

int scmgan_plane_colsum(const void* plane, int Cs, int c_off, int n, int B, int H, int W, float* S, float* db,
                        scmgan_stream_t stream) {
    SCM_REQUIRE(plane && (S || db), "colsum: null pointer");
    SCM_REQUIRE(n % 8 == 0 && n > 0 && n <= 256 && Cs % 8 == 0 && c_off % 8 == 0 && c_off + n <= Cs,
                "colsum: bad channel window");
    const int threads = 256;
    const int groups = n / 8;
    const int lanes = threads / groups;
    const int hw = H * W;
    int chunks = std::max(1, std::min((hw + 255) / 256, std::max(1, 4 * num_sms() / B)));
    const int rows_per_block = (hw + chunks - 1) / chunks;
    chunks = (hw + rows_per_block - 1) / rows_per_block;
    launch_k(plane_colsum_kernel, dim3(dim3(chunks, B)), dim3(threads), size_t(lanes * n * sizeof(float)), (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(plane), Cs, c_off, n, B, H, W, S, db, rows_per_block);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_spectral_norm_fwd(int count, const scmgan_sn_layer* layers, scmgan_stream_t stream) {
    return scmgan_spectral_norm_fwd_n(count, layers, 1, 0, stream);
}

int scmgan_spectral_norm_fwd_n(int count, const scmgan_sn_layer* layers, int iters, int sigma_stride,
                               scmgan_stream_t stream) {
    SCM_REQUIRE(count > 0 && count <= kMaxSnLayers && layers, "spectral_norm_fwd: bad layer count %d", count);
    SCM_REQUIRE(iters >= 1 && (iters == 1 || sigma_stride > 0), "spectral_norm_fwd: bad iteration count / sigma stride");
    SnLayers L;
    memset(&L, 0, sizeof(L));
    L.count = count;
    int max_smem = 0;
    for (int i = 0; i < count; ++i) {
        const scmgan_sn_layer& s = layers[i];
        SCM_REQUIRE(s.w && s.u && s.v && s.sigma && s.rows > 0 && s.cols > 0, "spectral_norm_fwd: bad layer %d", i);
        L.layer[i] = SnLayer{s.w, s.u, s.v, s.sigma, s.u_save, s.v_save, s.rows, s.cols};
        max_smem = std::max(max_smem, int((s.rows + s.cols + 64) * sizeof(float)));
    }
    // cluster kernel when every layer's column slice fits one 1024-thread pass
    {
        SnClusterGeom Gm;
        memset(&Gm, 0, sizeof(Gm));
        bool ok = true;
        static const char* off = getenv("SCMGAN_SN_SINGLE_CTA");
        if (off && atoi(off)) ok = false;
        int gmax = 1;
        for (int i = 0; i < count && ok; ++i) {
            const int cpc = (layers[i].cols + kSnCluster - 1) / kSnCluster;
            const int cpcp = (cpc + 31) & ~31;
            if (cpcp > 1024 || layers[i].rows > 1024) { ok = false; break; }
            Gm.cpc[i] = cpc;
            Gm.groups[i] = std::max(1, std::min(std::min(16, layers[i].rows), 1024 / cpcp));
            gmax = std::max(gmax, Gm.groups[i]);
            Gm.cpc_max = std::max(Gm.cpc_max, cpc);
            Gm.rows_max = std::max(Gm.rows_max, layers[i].rows);
        }
        if (ok) {
            const int smem = int(sizeof(float)) * (Gm.rows_max * (2 + kSnCluster) + Gm.cpc_max * (1 + gmax) +
                                                   kSnCluster + 33);
            if (smem <= 48 * 1024) {
                launch_k(sn_power_iter_cluster_kernel, dim3(count * kSnCluster), dim3(1024), size_t(smem), (cudaStream_t)stream, L, Gm, iters, sigma_stride);
                SCM_CUDA(cudaGetLastError());
                ++g_launches;
                return SCM_OK;
            }
        }
    }
    SCM_REQUIRE(max_smem <= 48 * 1024, "spectral_norm_fwd: layer too large");
    for (int it = 0; it < iters; ++it) {  // single-CTA fallback: one launch per iteration
        SnLayers Li = L;
        for (int i = 0; i < count; ++i) Li.layer[i].sigma = L.layer[i].sigma + (long long)it * sigma_stride;
        launch_k(sn_power_iter_kernel, dim3(count), dim3(1024), size_t(max_smem), (cudaStream_t)stream, Li);
        SCM_CUDA(cudaGetLastError());
        ++g_launches;
    }
    return SCM_OK;
}

int scmgan_spectral_norm_bwd(int count, const scmgan_sn_bwd_layer* layers, scmgan_stream_t stream) {
    SCM_REQUIRE(count > 0 && count <= kMaxSnLayers && layers, "spectral_norm_bwd: bad layer count %d", count);
    SnBwdLayers L;
    memset(&L, 0, sizeof(L));
    L.count = count;
    for (int i = 0; i < count; ++i) {
        const scmgan_sn_bwd_layer& s = layers[i];
        SCM_REQUIRE(s.g && s.wbar && s.u && s.v && s.sigma && s.dot && s.out, "spectral_norm_bwd: bad layer %d", i);
        L.layer[i] = SnBwdLayer{s.g, s.wbar, s.u, s.v, s.sigma, s.dot, s.out, s.rows, s.cols, s.accumulate, s.sigma2};
    }
    launch_k(sn_bwd_dot_kernel, dim3(dim3(96, count)), dim3(256), size_t(0), (cudaStream_t)stream, L);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    launch_k(sn_bwd_apply_kernel, dim3(dim3(256, count)), dim3(256), size_t(0), (cudaStream_t)stream, L);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_action_bias(const float* wbar, const float* sigma, const float* bias, const float* act, int B, int Cout,
                       int L, int A, float* out, scmgan_stream_t stream) {
    SCM_REQUIRE(wbar && act && out && B > 0 && Cout > 0 && A > 0, "action_bias: bad arguments");
    const int total = B * Cout;
    launch_k(action_bias_kernel, dim3((total + 127) / 128), dim3(128), size_t(0), (cudaStream_t)stream, wbar, sigma, bias, act, B, Cout, L, A, out);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_action_wgrad(const float* S, const float* act, int B, int Cout, int L, int A, float* g,
                        scmgan_stream_t stream) {
    SCM_REQUIRE(S && act && g && B > 0 && Cout > 0 && A > 0, "action_wgrad: bad arguments");
    const int total = Cout * A;
    launch_k(action_wgrad_kernel, dim3((total + 127) / 128), dim3(128), size_t(0), (cudaStream_t)stream, S, act, B, Cout, L, A, g);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_masked_mse_seq(const float* pred, const float* target, long long target_bstride, long long target_tstride,
                          const float* mask, long long mask_bstride, long long mask_tstride, int T, int B, int R,
                          float scale, const float* scale_dev, float* loss, float* loss_raw, float* dpred,
                          scmgan_stream_t stream) {
    SCM_REQUIRE(pred && target && loss && T > 0 && B > 0 && R > 0, "masked_mse: bad arguments");
    launch_k(masked_mse_kernel, dim3(1), dim3(256), size_t(0), (cudaStream_t)stream, pred, target, target_bstride, target_tstride, mask, mask_bstride, mask_tstride, T, B, R, scale, scale_dev, loss, loss_raw, dpred);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_masked_mse(const float* pred, const float* target, long long target_bstride, const float* mask,
                      long long mask_stride, int B, int R, float scale, const float* scale_dev, float* loss,
                      float* loss_raw, float* dpred, scmgan_stream_t stream) {
    return scmgan_masked_mse_seq(pred, target, target_bstride, 0, mask, mask_stride, 0, 1, B, R, scale, scale_dev, loss,
                                 loss_raw, dpred, stream);
}

int scmgan_bce_logits_seq(const float* x, const float* y, long long y_bstride, long long y_tstride, const float* mask,
                          long long mask_bstride, long long mask_tstride, int T, int B, long long per, float* loss_t,
                          float* dx, scmgan_stream_t stream) {
    SCM_REQUIRE(x && y && loss_t && T > 0 && B > 0 && per > 0, "bce_logits: bad arguments");
    SCM_REQUIRE((long long)T * B <= 65535, "bce_logits: T*B = %lld rows exceed the grid limit", (long long)T * B);
    const int threads = 256;
    int bx = int(std::min<long long>((per + threads * 4 - 1) / (threads * 4), std::max(1, 8 * num_sms() / (T * B))));
    bx = std::max(bx, 1);
    launch_k(bce_logits_kernel, dim3(dim3(bx, T * B)), dim3(threads), size_t(0), (cudaStream_t)stream, x, y, y_bstride, y_tstride, mask, mask_bstride, mask_tstride, T, B, per, loss_t, dx);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_bce_logits(const float* x, const float* y, long long y_bstride, const float* mask, int B, long long per,
                      float* loss, float* dx, scmgan_stream_t stream) {
    return scmgan_bce_logits_seq(x, y, y_bstride, 0, mask, 1, 0, 1, B, per, loss, dx, stream);
}

int scmgan_coord_wgrad(const float* dy, int B, int Co, int H, int W, float* g, long long g_s_co, long long g_s_ci,
                       int coord_c, scmgan_stream_t stream) {
    SCM_REQUIRE(dy && g && B > 0 && Co > 0 && H > 0 && W > 0 && coord_c >= 0, "coord_wgrad: bad arguments");
    launch_k(coord_wgrad_kernel, dim3(Co), dim3(256), size_t(0), (cudaStream_t)stream, dy, B, Co, H, W, g, g_s_co, g_s_ci, coord_c);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_eval_sqerr(const float* x, const float* y, long long y_bstride, long long y_tstride, int T, int B,
                      long long per, float* out, scmgan_stream_t stream) {
    SCM_REQUIRE(x && y && out && T > 0 && B > 0 && per > 0, "eval_sqerr: bad arguments");
    launch_k(eval_sqerr_kernel, dim3((unsigned)((long long)T * B)), dim3(256), size_t(0), (cudaStream_t)stream, x, y, y_bstride, y_tstride, B, per, out);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_eval_stats(const float* sqerr, const float* rpred, const float* rewards, long long r_bstride,
                      long long r_tstride, const float* dones, long long d_bstride, long long d_tstride, int T, int B,
                      int R, float* table, scmgan_stream_t stream) {
    SCM_REQUIRE(sqerr && rpred && rewards && dones && table && T > 0 && B > 0 && R > 0, "eval_stats: bad arguments");
    const size_t smem = (size_t(3) * B + 33) * sizeof(float);
    SCM_REQUIRE(smem <= 48 * 1024, "eval_stats: batch %d exceeds the single-block limit", B);
    launch_k(eval_stats_kernel, dim3(1), dim3(256), smem, (cudaStream_t)stream, sqerr, rpred, rewards, r_bstride, r_tstride, dones, d_bstride, d_tstride, T, B, R, table);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_reward_head_fwd(const float* y2, int B, int R, int H, int W, float* r, float* map,
                           scmgan_stream_t stream) {
    SCM_REQUIRE(y2 && r && B > 0 && R > 0 && 3 * R <= 16 && H >= 5 && W >= 5, "reward_head_fwd: bad arguments");
    const int h2 = (H - 5) / 2 + 1, w2 = (W - 5) / 2 + 1;
    launch_k(reward_head_fwd_kernel, dim3(dim3(B, R)), dim3(128), size_t(0), (cudaStream_t)stream, y2, B, R, H, W, h2, w2, r, map);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_reward_head_bwd(const float* y2, const float* dr, int B, int R, int H, int W, void* d2_plane,
                           scmgan_stream_t stream) {
    SCM_REQUIRE(y2 && dr && d2_plane && B > 0 && R > 0 && 3 * R <= 16 && H >= 5 && W >= 5,
                "reward_head_bwd: bad arguments");
    const int h2 = (H - 5) / 2 + 1, w2 = (W - 5) / 2 + 1;
    const long long rows = (long long)B * (H + 2) * (W + 2);
    launch_k(reward_head_bwd_kernel, dim3((unsigned)((rows + 127) / 128)), dim3(128), size_t(0), (cudaStream_t)stream, y2, dr, B, R, H, W, h2, w2, reinterpret_cast<__nv_bfloat16*>(d2_plane));
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_cf_loss_fwd(const float* za, const float* zb, const float* unswapped, const float* mask, int B, int L,
                       int HW, int mode, float lambda, float* rowmean, float* loss, scmgan_stream_t stream) {
    SCM_REQUIRE(za && zb && mask && rowmean && loss && B > 0 && L > 0 && L <= 64 && HW > 0, "cf_loss_fwd: bad arguments");
    SCM_REQUIRE(mode == 1 || (mode == 0 && unswapped), "cf_loss_fwd: mode 0 needs the unswapped-factor map");
    launch_k(cf_rowmean_kernel, dim3(B * L), dim3(256), size_t(0), (cudaStream_t)stream, za, zb, HW, rowmean);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    launch_k(cf_loss_fwd_kernel, dim3(1), dim3(256), size_t(0), (cudaStream_t)stream, rowmean, unswapped, mask, B, L, mode, lambda, loss);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_cf_loss_bwd(const float* za, const float* zb, const float* unswapped, const float* mask,
                       const float* rowmean, const float* gscale, int B, int L, int HW, int mode, float lambda,
                       float* dza, float* dzb, scmgan_stream_t stream) {
    SCM_REQUIRE(za && zb && mask && rowmean && gscale && (dza || dzb) && B > 0 && L > 0 && HW > 0,
                "cf_loss_bwd: bad arguments");
    const int bx = std::max(1, std::min((HW + 255) / 256, 8));
    launch_k(cf_loss_bwd_kernel, dim3(dim3(bx, B, L)), dim3(256), size_t(0), (cudaStream_t)stream, za, zb, unswapped, mask, rowmean, gscale, B, L, HW, mode, lambda, dza, dzb);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_transition_tail(const float* x, const float* uniforms, long long n, float* p, float* z,
                           scmgan_stream_t stream) {
    SCM_REQUIRE(x && z && n > 0, "transition_tail: bad arguments");
    const int blocks = int(std::min<long long>((n + 255) / 256, 4LL * num_sms()));
    launch_k(transition_tail_kernel, dim3(blocks), dim3(256), size_t(0), (cudaStream_t)stream, x, uniforms, n, p, z);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_decoder_bce_fwd(const scmgan_decoder_bce_desc* d, scmgan_stream_t stream) {
    SCM_REQUIRE(d != nullptr, "decoder_bce_fwd: null descriptor");
    const scmgan_conv_desc& c = d->conv;
    SCM_REQUIRE(d->target && d->loss_t && d->T > 0 && d->B > 0 && c.B == d->T * d->B,
                "decoder_bce_fwd: bad arguments (conv.B must be T*B, t-major)");
    SCM_REQUIRE(c.n == 16 && c.out && c.out_c_off == 0 && c.act == SCMGAN_ACT_NONE && !c.gate && !c.add && !c.wrap &&
                !c.sample_out, "decoder_bce_fwd: conv2 must be a plain 16-column head writing a zero-halo plane");
    SCM_REQUIRE(c.n_valid > 0 && c.n_valid <= 16, "decoder_bce_fwd: bad n_valid=%d", c.n_valid);
    // per-warp partial sums: one row [T] per epilogue warp of every CTA the weight-stationary kernel can launch
    const int ws_rows = scmgan_decoder_bce_workspace_rows();
    SCM_REQUIRE(d->workspace && d->workspace_bytes >= (long long)ws_rows * d->T * 4,
                "decoder_bce_fwd: workspace of %d x T floats required", ws_rows);
    cudaStream_t st = (cudaStream_t)stream;
    // rows of CTAs / warps that do not exist in this launch must read as zero
    SCM_CUDA(cudaMemsetAsync(d->workspace, 0, size_t(ws_rows) * d->T * 4, st));
    t_bce = d;
    const int rc = conv_impl(&c, st);
    t_bce = nullptr;
    if (rc) return rc;
    launch_k(bce_finalize_kernel, dim3(d->T), dim3(256), size_t(0), st, (const float*)d->workspace, ws_rows, d->T, d->loss_t);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_decoder_bce_workspace_rows(void) { return 2 * num_sms() * 16; }

int scmgan_decoder_bce_bwd(void* dlogits_plane, int cs, int fmt, const float* g, int T, int B, int H, int W,
                           scmgan_stream_t stream) {
    SCM_REQUIRE(dlogits_plane && g && cs > 0 && cs % 8 == 0 && T > 0 && B > 0 && H > 0 && W > 0 && (fmt | 1) == 1,
                "decoder_bce_bwd: bad arguments");
    const long long per_step = (long long)B * (H + 2) * (W + 2) * cs;  // 16-bit elements of one rollout step
    SCM_REQUIRE(per_step % 8 == 0, "decoder_bce_bwd: plane size");
    const int bx = int(std::min<long long>((per_step / 8 + 255) / 256, std::max(1, 8 * num_sms() / T)));
    launch_k(plane_scale_steps_kernel, dim3(dim3(bx, T)), dim3(256), size_t(0), (cudaStream_t)stream,
             reinterpret_cast<uint4*>(dlogits_plane), per_step / 8, g, fmt);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

static int csrn_fill(const scmgan_csrn_sweep_desc* d, CsrnSweepParams& P, bool bwd) {
    SCM_REQUIRE(d != nullptr, "gru_conv_sweep: null descriptor");
    SCM_REQUIRE(d->x && d->w_ih && d->w_hh && d->conv_w && d->conv_b && d->ctx, "gru_conv_sweep: null pointer");
    SCM_REQUIRE(d->B > 0 && d->C > 0 && d->L > 0 && d->n > 0, "gru_conv_sweep: bad geometry");
    SCM_REQUIRE(!bwd || (d->states && d->dctx && d->dx && d->dparams), "gru_conv_sweep_bwd: null pointer");
    memset(&P, 0, sizeof(P));
    P.x = d->x; P.xs_b = d->xs_b; P.xs_c = d->xs_c; P.xs_line = d->xs_line; P.xs_pix = d->xs_pix;
    P.B = d->B; P.C = d->C; P.L = d->L; P.n = d->n; P.reverse = d->reverse;
    P.w_ih = d->w_ih; P.w_hh = d->w_hh; P.conv_w = d->conv_w; P.conv_b = d->conv_b;
    P.states = d->states;
    if (bwd) {
        P.ctx = const_cast<float*>(d->dctx); P.ctx_fwd = d->ctx; P.dx = d->dx; P.dparams = d->dparams;
    } else {
        P.ctx = d->ctx;
    }
    return SCM_OK;
}

int scmgan_gru_conv_sweep_fwd(const scmgan_csrn_sweep_desc* d, scmgan_stream_t stream) {
    CsrnSweepParams P;
    int rc = csrn_fill(d, P, false);
    if (rc) return rc;
    const size_t smem = size_t(3) * P.n * P.C * sizeof(float);
    if (smem > size_t(kSmemMax)) {
        set_error("gru_conv_sweep_fwd: line of %d x %d does not fit shared memory", P.n, P.C);
        return SCM_EUNSUPPORTED;
    }
    SCM_OPT_IN_SMEM(csrn_sweep_fwd_kernel, kSmemMax);
    launch_k(csrn_sweep_fwd_kernel, dim3(P.B), dim3(256), size_t(smem), (cudaStream_t)stream, P);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_gru_conv_sweep_bwd(const scmgan_csrn_sweep_desc* d, scmgan_stream_t stream) {
    CsrnSweepParams P;
    int rc = csrn_fill(d, P, true);
    if (rc) return rc;
    const size_t smem = size_t(12) * P.n * P.C * sizeof(float);
    if (smem > size_t(kSmemMax)) {
        set_error("gru_conv_sweep_bwd: line of %d x %d does not fit shared memory", P.n, P.C);
        return SCM_EUNSUPPORTED;
    }
    SCM_OPT_IN_SMEM(csrn_sweep_bwd_kernel, kSmemMax);
    launch_k(csrn_sweep_bwd_kernel, dim3(P.B), dim3(256), size_t(smem), (cudaStream_t)stream, P);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_philox_uniform(float* out, long long n, unsigned long long* rng_state, scmgan_stream_t stream) {
    SCM_REQUIRE(out && rng_state && n > 0, "philox_uniform: bad arguments");
    const long long blocks4 = (n + 3) / 4;
    launch_k(philox_fill_kernel, dim3(unsigned((blocks4 + 255) / 256)), dim3(256), size_t(0), (cudaStream_t)stream, out, n, rng_state);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    launch_k(rng_advance_kernel, dim3(1), dim3(1), size_t(0), (cudaStream_t)stream, rng_state, (unsigned long long)blocks4);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_replay_sample(const scmgan_replay_desc* d, scmgan_stream_t stream) {
    SCM_REQUIRE(d != nullptr, "replay_sample: null descriptor");
    SCM_REQUIRE(d->frames && d->rewards && d->actions && d->ep_len && d->n_filled && d->rng_state,
                "replay_sample: null buffer pointer");
    SCM_REQUIRE(d->states && d->rewards_out && d->dones && d->actions_out, "replay_sample: null output pointer");
    SCM_REQUIRE(d->slots > 0 && d->max_len >= 4 && d->per_frame > 0 && d->R > 0 && d->B > 0 && d->Hn > 0,
                "replay_sample: bad geometry");
    SCM_REQUIRE(size_t(3) * d->Hn * sizeof(int) <= 48 * 1024, "replay_sample: %d timesteps exceed the row buffer", d->Hn);
    ReplayParams P;
    P.frames = d->frames; P.rewards = d->rewards; P.actions = d->actions; P.ep_len = d->ep_len;
    P.n_filled = d->n_filled; P.slots = d->slots; P.max_len = d->max_len; P.R = d->R; P.per_frame = d->per_frame;
    P.B = d->B; P.Hn = d->Hn; P.random_start = d->random_start; P.rng = d->rng_state;
    P.states = d->states; P.rewards_out = d->rewards_out; P.dones = d->dones; P.actions_out = d->actions_out;
    P.plan = d->plan;
    launch_k(replay_sample_kernel, dim3(d->B), dim3(256), size_t(3) * d->Hn * sizeof(int), (cudaStream_t)stream, P);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    launch_k(rng_advance_kernel, dim3(1), dim3(1), size_t(0), (cudaStream_t)stream, d->rng_state, (unsigned long long)d->B * d->Hn);
    SCM_CUDA(cudaGetLastError());
    ++g_launches;
    return SCM_OK;
}

int scmgan_clip_adam(int count, const scmgan_adam_chunk* chunks, float lr, float beta1, float beta2, float eps,
                     int step, const float* step_dev, float gscale, scmgan_stream_t stream) {
    SCM_REQUIRE(count >= 0 && (count == 0 || chunks), "clip_adam: bad arguments");
    SCM_REQUIRE(step_dev || step >= 1, "clip_adam: step must be >= 1");
    for (int base = 0; base < count; base += kAdamChunksPerLaunch) {
        AdamArgs A;
        memset(&A, 0, sizeof(A));
        A.count = std::min(kAdamChunksPerLaunch, count - base);
        int max_n = 0;
        for (int i = 0; i < A.count; ++i) {
            const scmgan_adam_chunk& c = chunks[base + i];
            SCM_REQUIRE(c.p && c.g && c.m && c.v && c.n > 0, "clip_adam: bad chunk %d", base + i);
            A.chunk[i] = AdamChunk{c.p, c.g, c.m, c.v, c.n, c.clip, c.step};
            max_n = std::max(max_n, c.n);
        }
        A.lr = lr; A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.step_ptr = step_dev; A.gscale = gscale;
        if (!step_dev) {
            A.bc1 = 1.f - powf(beta1, float(step));
            A.bc2_sqrt = sqrtf(1.f - powf(beta2, float(step)));
        }
        const int bx = std::max(1, std::min((max_n + 1023) / 1024, 64));
        launch_k(clip_adam_kernel, dim3(dim3(bx, A.count)), dim3(256), size_t(0), (cudaStream_t)stream, A);
        SCM_CUDA(cudaGetLastError());
    ++g_launches;
    }
    return SCM_OK;
}

}  // extern "C"
