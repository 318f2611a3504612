// 3x3 convolution with 16 input channels ("expand" layers: the first conv of every network and its mirror in the
// backward pass - reference models.py:51,129,237,260 and the data gradients of conv6 / conv4 / the last decoder and
// reward convs): K = 9 x 16 = 144, N = 64 or 128 output channels.  Nine MMAs per 128-pixel tile - the layer is a
// pure store stream (2 N bytes written per pixel against 32 read), so the kernel is built around its epilogue:
//
//  * Tiles are IMAGE-ALIGNED: k = floor(128 / W) whole rows of one sample (128 pixels = 2 rows at W = 64), interior
//    pixels only.  The flat-plane tiles of conv_igemm*.cuh contain halo rows whose accumulators are garbage, which
//    rules out bulk stores; an image-aligned tile is a dense [k][W] box of the output tensor.
//  * That needs an A operand whose 128 rows are the pixels (r, w + kx - 1) of k consecutive image rows - not a run of
//    the flattened plane.  With 32-byte pixel rows the producer warps gather it themselves (cp.async, 16 bytes per
//    request): one dense copy [k + 2][W] of the input per horizontal tap kx, so that tap (ky, kx) is the 128
//    consecutive rows starting at row ky * W of copy kx.  (A TMA box of 32-byte rows is bound by the TMA request rate:
//    measured 3.7k cycles per 36 KB in round 1.)  The manual 32-byte swizzle (address bit 4 ^= bit 7 of the absolute
//    shared-memory address) is what the tensor core applies to the descriptors, see conv_igemm_v2.cuh.
//  * Epilogue: TMEM -> registers -> bias / per-sample bias / LeakyReLU / LeakyReLU-derivative gate -> 16-bit ->
//    shared-memory staging tile in the 128-byte-swizzled layout -> ONE TMA STORE per 64-channel half and tile
//    (cp.async.bulk.tensor ... global.shared::cta through an interior-view tensor map: rows past the image end are
//    clipped by the hardware).  Two staging buffers per half, so the store of tile i drains while tile i + 1 is
//    processed.  The gate plane of the gated variant arrives the same way in the other direction (TMA load of the
//    tile's [k][W][64] box per half), instead of 256-byte per-lane loads.
//  * Only the halo copies (wrapped border of Transition planes, zero border of the others: the pixels of the first /
//    last row and column, 6 % at 64 x 64) are written by per-lane stores.
//  * Sixteen epilogue warps: lane quarter x channel half x 32-column pass.  Measured (in-kernel clock64 timeline,
//    profiles/expand_timeline.py): a lone warp per scheduler retires a dependent instruction every ~7 cycles, a
//    32-column pass takes ~1000 cycles whatever it contains - only more resident warps hide that.  Each channel half is
//    an independent store pipeline (own named barrier, own TMA store group); per-tile bias vectors sit in warp-private
//    shared memory; tile coordinates advance incrementally (an integer division costs ~100 cycles here).
//  * Four producer warps with incremental addressing: the first version spent 4.2k cycles per tile on the address
//    arithmetic of its 24 copies per lane and was the bottleneck of the whole kernel.
#pragma once
#include "conv_igemm.cuh"

namespace scm {

struct ExpandParams {
    int B, H, W, Hp, Wp;
    int k;              // image rows per tile
    int tiles_per_img;  // ceil(H / k)
    int num_tiles;      // B * tiles_per_img
    int n;              // output channels of the GEMM: 64 or 128
    const __nv_bfloat16* a;  // input plane [B][Hp][Wp][a_cs], 16 channels from a_c_off
    int a_cs, a_c_off;
    int copy_rows;      // (k + 2) * W rows of one horizontal-tap copy
    int copy_bytes;     // its size rounded up to 1024
    int num_a_stages;
    float scale;
    const float* bias;         // [bias_n] or nullptr
    int bias_n;
    const float* sample_bias;  // [B][n] or nullptr (replaces bias)
    const float* sample_scale; // [B] or nullptr
    int act;
    float slope;
    __nv_bfloat16* out;  // output plane (for the per-lane halo copies; the interior goes through tmap_out)
    int out_cs, out_c_off;
    int wrap;            // 1: halo = wrapped border, 0: halo = zeros
    int gated;           // multiply by lrelu'(gate) (gate tile through tmap_gate)
    int gate_c_off;
    int a_fmt, b_fmt, out_fmt;
    int debug;  // profiling aid (env SCMGAN_DEBUG): 1 no staging/TMA store, 2 no TMEM loads, 4 no MMAs, 8 no A copies,
                // 16 no bias loads, 32 no epilogue arithmetic, 64 no halo copies
};

// In-kernel timeline (profiling aid, SCMGAN_DEBUG bit 4096; read back by profiles/expand_timeline.py): clock64 stamps
// of CTA 0, [role: producer warp 0 / MMA / first epilogue warp][tile iteration][event].
__device__ unsigned long long g_exp_dbg[3 * 16 * 8];
#define EXP_STAMP(role, it, ev)                                                                     \
    do {                                                                                            \
        if ((P.debug & 4096) && blockIdx.x == 0 && lane == 0 && (it) < 16)                          \
            g_exp_dbg[((role) * 16 + (it)) * 8 + (ev)] = (unsigned long long)clock64();             \
    } while (0)

constexpr int kExpEpiWarps = 16;
constexpr int kExpProdWarps = 4;
constexpr int kExpThreads = 32 * (kExpProdWarps + 2 + kExpEpiWarps);  // 4 producers, MMA, gate loader, 16 epilogue
constexpr int kExpAccStages = 4;                                      // 4 x 128 TMEM columns

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    tmem_ld16(taddr, v);
    tmem_ld16(taddr + 16, v + 16);
}

// tile -> (sample b, first output row h0) without divisions: advance by the grid stride with a carry
struct ExpTileIter {
    int tile, b, ti, dq, dr, tpi;
    __device__ __forceinline__ ExpTileIter(int first, int stride, int tiles_per_img) {
        tile = first; tpi = tiles_per_img;
        b = first / tiles_per_img; ti = first - b * tiles_per_img;
        dq = stride / tiles_per_img; dr = stride - dq * tiles_per_img;
    }
    __device__ __forceinline__ void next(int stride) {
        tile += stride; b += dq; ti += dr;
        if (ti >= tpi) { ti -= tpi; ++b; }
    }
};

__global__ void __launch_bounds__(kExpThreads, 1)
conv3x3_expand_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
                      const __grid_constant__ CUtensorMap tmap_gate, const __grid_constant__ ExpandParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int halves = P.n / 64;
    const int b_tile_bytes = P.n * 32;                       // one tap: [n][16] K-major, 32B swizzle
    const int b_bytes = (9 * b_tile_bytes + 1023) & ~1023;
    const int a_stage_bytes = 3 * P.copy_bytes;
    constexpr int kStageTile = 128 * 128;                    // staging / gate tile of one half: 128 rows x 128 B
    uint8_t* s_b = smem;
    uint8_t* s_a = s_b + b_bytes;
    uint8_t* s_out = s_a + size_t(P.num_a_stages) * a_stage_bytes;   // [half][2][16 KB]
    uint8_t* s_gate = s_out + size_t(halves) * 2 * kStageTile;       // [2][half][16 KB] (gated only)
    uint8_t* s_tail = s_gate + (P.gated ? size_t(halves) * 2 * kStageTile : 0);
    float* s_bias = reinterpret_cast<float*>(s_tail);                // [16 warps][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kExpEpiWarps * 32);
    uint64_t* a_full = bars;             // [4]
    uint64_t* a_empty = bars + 4;        // [4]
    uint64_t* acc_full = bars + 8;       // [4]
    uint64_t* acc_empty = bars + 12;     // [4]
    uint64_t* g_full = bars + 16;        // [2]
    uint64_t* g_empty = bars + 18;       // [2]
    uint64_t* b_full = bars + 20;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_out);
        if (P.gated) prefetch_tmap(&tmap_gate);
        for (int s = 0; s < 4; ++s) {
            mbar_init(&a_full[s], 32 * kExpProdWarps);  // one cp.async arrival per producer lane
            mbar_init(&a_empty[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 8 * halves);       // one arrival per epilogue warp
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&g_full[s], 1);
            mbar_init(&g_empty[s], 8 * halves);
        }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == kExpProdWarps) {  // the MMA warp owns the TMEM allocation
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    pdl_sync();  // everything above is CTA-local: it overlaps the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kExpProdWarps) {
        // ------------------------------ A producers: three dense shifted copies per tile ------------------------------
        // Lane l of producer warp pw owns 16-byte chunk c = pw*32 + l (+128 j) of every image row: pixel c>>1, half c&1.
        const int c0 = warp * 32 + lane;
        const int row_chunks = 2 * P.W;
        const uint32_t pix_bytes = uint32_t(P.a_cs) * 2u;                 // plane pixel pitch
        const uint32_t src_lane = uint32_t(c0 >> 1) * pix_bytes + uint32_t(c0 & 1) * 16u;
        const uint32_t src_step = 64u * pix_bytes;                         // 128 chunks = 64 pixels further
        const uint64_t img_bytes = uint64_t(P.Hp) * P.Wp * pix_bytes;
        const uint32_t src_row_bytes = uint32_t(P.Wp) * pix_bytes;
        const char* plane0 = reinterpret_cast<const char*>(P.a + P.a_c_off);
        const int rows = P.k + 2;
        int stage = 0;
        uint32_t phase = 0;
        ExpTileIter it(blockIdx.x, stride, P.tiles_per_img);
        int n_it = 0;
        for (; it.tile < P.num_tiles; it.next(stride), ++n_it) {
            const int h0 = it.ti * P.k;  // first output row (interior coordinates) = padded row of the tap row ky = 0
            if (warp == 0) EXP_STAMP(0, n_it, 0);
            mbar_wait(&a_empty[stage], phase ^ 1);
            if (warp == 0) EXP_STAMP(0, n_it, 1);
            if (!(P.debug & 8)) {
                const uint32_t sa = smem_u32(s_a + size_t(stage) * a_stage_bytes);
                const char* src_t = plane0 + uint64_t(it.b) * img_bytes + uint64_t(h0) * src_row_bytes + src_lane;
                const int rows_ok = min(rows, P.Hp - h0);  // rows past the plane end are zero-filled
                uint32_t dst_k = sa + uint32_t(c0) * 16u;
                for (int kx = 0; kx < 3; ++kx) {
                    const char* src_r = src_t + uint32_t(kx) * pix_bytes;
                    uint32_t dst_r = dst_k;
                    for (int r = 0; r < rows; ++r) {
                        const uint32_t nbytes = r < rows_ok ? 16u : 0u;
                        const char* src = nbytes ? src_r : plane0;
                        uint32_t dst = dst_r;
                        for (int c = c0; c < row_chunks; c += 128) {
                            // 32-byte swizzle (smem address bit 4 ^= bit 7), applied to the SOURCE: the lane that writes
                            // smem chunk `dst` fetches the other 16-byte half of its pixel when bit 7 of dst is set, so
                            // that the shared-memory addresses of a warp stay linear in lane order
                            const uint32_t flip = (dst >> 3) & 16u;
                            if (P.debug & 8192) {  // profiling: previous scheme (destination permuted)
                                const uint32_t ad = dst ^ flip;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ad), "l"(src), "r"(nbytes)
                                             : "memory");
                            } else {
                                const char* s2 = (c & 1) ? src - flip : src + flip;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(s2), "r"(nbytes)
                                             : "memory");
                            }
                            src += nbytes ? src_step : 0u;
                            dst += 128u * 16u;
                        }
                        src_r += src_row_bytes;
                        dst_r += uint32_t(row_chunks) * 16u;
                    }
                    dst_k += uint32_t(P.copy_bytes);
                }
            }
            if (warp == 0) EXP_STAMP(0, n_it, 2);
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&a_full[stage])) : "memory");
            if (++stage == P.num_a_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == kExpProdWarps) {
        // ------------------------------ MMA issuer ------------------------------
        if (elect_one()) {
            mbar_arrive_expect_tx(b_full, uint32_t(9 * b_tile_bytes));
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_b + tap * b_tile_bytes, &tmap_b, b_full, 0, tap * P.n);
        }
        __syncwarp();
        const uint32_t idesc = make_idesc_ab(128, P.n, P.a_fmt, P.b_fmt, 0, 0);
        const uint64_t bdesc0 = make_smem_desc(smem_u32(s_b), 16, 256, kLayoutSw32);
        const uint64_t adesc0 = make_smem_desc(smem_u32(s_a), 16, 256, kLayoutSw32);
        const uint32_t b_tile16 = uint32_t(b_tile_bytes) >> 4;
        const uint32_t a_stage16 = uint32_t(a_stage_bytes) >> 4;
        const uint32_t copy16 = uint32_t(P.copy_bytes) >> 4;
        const uint32_t row16 = uint32_t(P.W) * 2u;  // one image row of 32-byte pixels, in 16-byte units
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        mbar_wait(b_full, 0);
        int n_it = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += stride, ++n_it) {
            EXP_STAMP(1, n_it, 0);
            mbar_wait(&acc_empty[acc], acc_phase ^ 1);
            EXP_STAMP(1, n_it, 1);
            mbar_wait(&a_full[stage], phase);
            EXP_STAMP(1, n_it, 2);
            fence_proxy_async_smem();  // cp.async (generic proxy) -> tensor core
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + uint32_t(acc * 128);
            const uint64_t a_st = adesc0 + uint64_t(uint32_t(stage) * a_stage16);
            if (elect_one()) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    if (P.debug & 4) break;
                    const int ky = tap / 3, kx = tap % 3;
                    umma_f16(tmem_d, a_st + uint64_t(uint32_t(kx) * copy16 + uint32_t(ky) * row16),
                             bdesc0 + uint64_t(uint32_t(tap) * b_tile16), idesc, tap > 0 ? 1u : 0u);
                }
                umma_commit(&a_empty[stage]);
                umma_commit(&acc_full[acc]);
            }
            __syncwarp();
            EXP_STAMP(1, n_it, 3);
            if (++stage == P.num_a_stages) { stage = 0; phase ^= 1; }
            if (++acc == kExpAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == kExpProdWarps + 1) {
        // ------------------------------ gate loader (gated variant): the tile's [k][W][64] box per half ------------------------------
        if (P.gated) {
            int g = 0;
            uint32_t gphase = 0;
            ExpTileIter it(blockIdx.x, stride, P.tiles_per_img);
            for (; it.tile < P.num_tiles; it.next(stride)) {
                mbar_wait(&g_empty[g], gphase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&g_full[g], uint32_t(halves * P.k * P.W * 128));
                    for (int hf = 0; hf < halves; ++hf)
                        tma_load_4d(s_gate + size_t(g * halves + hf) * kStageTile, &tmap_gate, &g_full[g],
                                    P.gate_c_off + hf * 64, 0, it.ti * P.k, it.b);
                }
                __syncwarp();
                if (++g == 2) { g = 0; gphase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue: lane quarter q x channel half hf x 32-column pass ------------------------------
        const int ew = warp - (kExpProdWarps + 2);
        const int q = warp & 3;          // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
        const int pass = (ew >> 2) & 1;  // which 32 of the half's 64 columns
        const int hf = ew >> 3;          // channel half
        if (hf < halves) {
            const int m = q * 32 + lane;  // tile row = r * W + w
            const int r = m / P.W;
            const int w = m - r * P.W;
            float* my_bias = s_bias + ew * 32;
            const uint32_t my_bias_u32 = smem_u32(my_bias);
            const int bar_id = 1 + hf;
            const bool issuer = (q == 0 && pass == 0 && lane == 0);
            const int col0 = hf * 64 + pass * 32;  // first GEMM column of this warp
            const uint32_t stg_row = uint32_t(m) * 128u;
            const uint32_t swz = uint32_t(m & 7);
            const bool per_sample = P.sample_bias != nullptr || P.sample_scale != nullptr;
            int acc = 0, g = 0, buf = 0, last_b = -1;
            uint32_t acc_phase = 0, gphase = 0;
            float rs = P.scale;
            ExpTileIter it(blockIdx.x, stride, P.tiles_per_img);
            int n_it = 0;
            for (; it.tile < P.num_tiles; it.next(stride), ++n_it) {
                if (ew == 0) EXP_STAMP(2, n_it, 0);
                const int b = it.b;
                const int h0 = it.ti * P.k;
                const int h = h0 + r;
                const bool valid = r < P.k && h < P.H;
                // this tile's bias vector (warp-private copy; refreshed when the sample changes)
                if ((last_b < 0 || (per_sample && b != last_b)) && !(P.debug & 16)) {
                    const int c = col0 + lane;
                    float bv;
                    if (P.sample_bias) bv = __ldg(P.sample_bias + size_t(b) * P.n + c);
                    else bv = (P.bias && c < P.bias_n) ? __ldg(P.bias + c) : 0.f;
                    if (P.sample_scale) rs = P.scale * __ldg(P.sample_scale + b);
                    __syncwarp();
                    my_bias[lane] = bv;
                    __syncwarp();
                    last_b = b;
                }
                uint8_t* stg = s_out + size_t(hf * 2 + buf) * kStageTile;
                // the staging buffer was handed to the TMA two tiles ago: its store must have finished reading
                if (issuer) tma_store_wait_read<1>();
                if (ew == 0) EXP_STAMP(2, n_it, 1);
                named_bar_sync(bar_id, 256);
                if (ew == 0) EXP_STAMP(2, n_it, 2);
                mbar_wait(&acc_full[acc], acc_phase);
                if (ew == 0) EXP_STAMP(2, n_it, 3);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 128 + col0);
                float v[32];
                if (!(P.debug & 2)) {
                    tmem_ld32(taddr, v);   // in flight while the gate bits are extracted
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = float(i + lane);
                }
                uint32_t gm = 0xFFFFFFFFu;
                if (P.gated) {
                    mbar_wait(&g_full[g], gphase);
                    const uint8_t* gt = s_gate + size_t(g * halves + hf) * kStageTile + stg_row;
                    gm = 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 gv = *reinterpret_cast<const uint4*>(gt + ((uint32_t(pass * 4 + j) ^ swz) << 4));
                        const uint32_t wv[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            // per 16-bit half: "> 0" <=> magnitude != 0 and sign clear (same bits in fp16 and bf16)
                            const uint32_t pos = ((wv[t] & 0x7FFF7FFFu) + 0x7FFF7FFFu) & ~wv[t] & 0x80008000u;
                            gm |= ((pos >> 15) & 1u) << (8 * j + 2 * t);
                            gm |= (pos >> 31) << (8 * j + 2 * t + 1);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&g_empty[g]);
                }
                if (!(P.debug & 2)) tmem_ld_wait();
                if (ew == 0) EXP_STAMP(2, n_it, 4);
                if (!(P.debug & 32)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b4;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                                     : "r"(my_bias_u32 + uint32_t(4 * j) * 4u));
                        v[4 * j] = fmaf(v[4 * j], rs, b4.x);
                        v[4 * j + 1] = fmaf(v[4 * j + 1], rs, b4.y);
                        v[4 * j + 2] = fmaf(v[4 * j + 2], rs, b4.z);
                        v[4 * j + 3] = fmaf(v[4 * j + 3], rs, b4.w);
                    }
                    if (P.act == ACT_LRELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], v[i] * P.slope);
                    }
                    if (P.gated) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= ((gm >> i) & 1u) ? 1.f : P.slope;
                    }
                }
                uint32_t o[16];
                if (P.out_fmt == FMT_F16) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = pack2_f16(v[2 * i], v[2 * i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = pack2_bf16(v[2 * i], v[2 * i + 1]);
                }
                if (ew == 0) EXP_STAMP(2, n_it, 5);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);   // the accumulator columns are in registers now
                // staging tile row m, 16-byte chunks pass*4 .. pass*4+3, 128-byte swizzle
                if (!(P.debug & 1)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(stg + stg_row + ((uint32_t(pass * 4 + j) ^ swz) << 4)) =
                            make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                }
                // halo copies of border pixels: wrapped value (Transition planes) or zeros
                if (valid && !(P.debug & 64) && (h == 0 || h == P.H - 1 || w == 0 || w == P.W - 1)) {
                    int hp2 = -1, wp2 = -1;
                    if (h == 0) hp2 = P.wrap ? P.H + 1 : 0; else if (h == P.H - 1) hp2 = P.wrap ? 0 : P.H + 1;
                    if (w == 0) wp2 = P.wrap ? P.W + 1 : 0; else if (w == P.W - 1) wp2 = P.wrap ? 0 : P.W + 1;
                    if (!P.wrap) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = 0u;
                    }
                    __nv_bfloat16* ob = P.out + size_t(b) * P.Hp * P.Wp * P.out_cs + P.out_c_off + col0;
                    const uint4 o0 = make_uint4(o[0], o[1], o[2], o[3]), o1 = make_uint4(o[4], o[5], o[6], o[7]);
                    const uint4 o2 = make_uint4(o[8], o[9], o[10], o[11]), o3 = make_uint4(o[12], o[13], o[14], o[15]);
                    auto put = [&](int hp, int wp) {
                        uint4* d = reinterpret_cast<uint4*>(ob + (size_t(hp) * P.Wp + wp) * P.out_cs);
                        d[0] = o0; d[1] = o1; d[2] = o2; d[3] = o3;
                    };
                    if (hp2 >= 0) put(hp2, w + 1);
                    if (wp2 >= 0) put(h + 1, wp2);
                    if (hp2 >= 0 && wp2 >= 0) put(hp2, wp2);
                }
                if (ew == 0) EXP_STAMP(2, n_it, 6);
                fence_proxy_async_smem();  // staging writes -> async proxy (TMA store)
                named_bar_sync(bar_id, 256);
                if (ew == 0) EXP_STAMP(2, n_it, 7);
                if (issuer && !(P.debug & 1)) {
                    tma_store_4d(&tmap_out, stg, P.out_c_off + hf * 64, 0, h0, b);
                    tma_store_commit();
                }
                buf ^= 1;
                if (++acc == kExpAccStages) { acc = 0; acc_phase ^= 1; }
                if (P.gated && ++g == 2) { g = 0; gphase ^= 1; }
            }
            if (issuer) tma_store_wait_all();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kExpProdWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace scm
