// Weight gradient of the 16-output-channel convolutions (the last conv of every network: reference models.py:55, 133,
// 233, 264): dW[co][ci][tap] = sum_p dY[p][co] * X[p + tap][ci] with only 16 values of co.
//
// As a GEMM per tap this is M = 128 (ci), N = 16: a tcgen05.mma with N = 16 costs 39 cycles against 64 for N = 128
// (profiles/mma_rate.cu - the 128 x 16 A block is re-read from shared memory for every instruction), and the nine taps
// each need their own shifted copy of the big operand X.  conv_wgrad.cuh ran it that way: 26 us + a 9 us reduction for
// 1/8 of the FLOPs of a 128 x 128 layer.
//
// Here the nine taps are part of N instead.  With q = p + tap,
//     dW[ci][(tap, co)] = sum_q X[q][ci] * dY[q - tap][co]
// so ONE MMA per 16 pixels has A = X (padded view, unshifted, loaded and read once) and B = nine 16-channel tiles of dY,
// each fetched through the interior-view tensor map at its tap's offset (TMA zero-fills what falls outside the H x W
// interior, which is exactly the pixels that tap does not see).  N = 9 x 16 = 144; B is MN-major with the 32-byte
// swizzle, one 16-channel atom per tap (LBO = tile size).  A K block is BH padded image rows of kw >= W + 2 pixels.
//
// The mirror case, 16 INPUT channels (the first conv of every network), is the same kernel with the roles swapped:
// A = dY (interior view, 128 channels of co), B = nine tiles of X through the padded view at +tap, and a tenth,
// constant all-ones tile whose output column is the bias gradient sum_p dY[p][co].
#pragma once
#include "conv_wgrad.cuh"

namespace scm {

struct WgradNarrowParams {
    int B, Hp;
    int BH;           // padded image rows per stage
    int nrb;          // ceil(Hp / BH)
    int num_kblocks;  // B * nrb
    int kb_per_cta;
    int kw;           // pixels per image-row box (multiple of 16, >= W + 2)
    int x_c_off, dy_c_off;  // channel offsets of the 128-wide (M) and the 16-wide (N) operand
    int n_sign;       // N-operand box origin = (n_sign*kx, h0 + n_sign*ky): -1 = dY seen from X, +1 = X seen from dY
    int with_ones;    // append the all-ones tile (bias gradient; only meaningful when M is dY)
    float* ws;        // split-K partials [m block][split][tap][16][128]
    float* ws_bias;   // [m block][split][128] (with_ones)
    int debug;
    int m_fmt, n_fmt;  // 16-bit formats of the 128-wide (M) and the 16-wide (N) operand planes
};

constexpr int kNarrowCo = 16;
constexpr int kNarrowN = 9 * kNarrowCo;

// grid = (splits, 128-channel blocks of ci)
__global__ void __launch_bounds__(kWgradThreads, 1)
conv3x3_wgrad_narrow_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                            const __grid_constant__ WgradNarrowParams P, int num_stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KP = P.kw * P.BH;             // pixels (GEMM K) per stage
    const int a_atom_bytes = KP * 128;      // 64 channels x KP pixels
    const int a_bytes = 2 * a_atom_bytes;
    const int b_tile_bytes = KP * 32;       // 16 channels x KP pixels, one per tap
    const int n_tiles = 9 + (P.with_ones ? 1 : 0);
    const int stage_bytes = (a_bytes + 10 * b_tile_bytes + 1023) & ~1023;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(num_stages) * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* acc_full = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    uint64_t* conv_bar = bars + 24;  // [8] fp16 operand of the stage rewritten to bf16 (conv_wgrad.cuh)
    const bool conv_m = P.m_fmt != P.n_fmt && P.m_fmt == FMT_F16;
    const bool conv_n = P.m_fmt != P.n_fmt && P.n_fmt == FMT_F16;
    const bool convert = conv_m || conv_n;
    const int mma_fmt = convert ? FMT_BF16 : P.m_fmt;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kb_begin = blockIdx.x * P.kb_per_cta;
    const int kb_end = min(P.num_kblocks, kb_begin + P.kb_per_cta);
    const int nkb = max(0, kb_end - kb_begin);
    const int x_c0 = P.x_c_off + blockIdx.y * 128;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dy);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&conv_bar[s], kWgradConvThreads);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    pdl_sync();
    if (P.with_ones) {  // tile 9 of every stage: bf16 1.0 (never touched by TMA)
        for (int s = 0; s < num_stages; ++s) {
            uint32_t* ones = reinterpret_cast<uint32_t*>(smem + size_t(s) * stage_bytes + a_bytes + 9 * b_tile_bytes);
            for (int i = threadIdx.x; i < b_tile_bytes / 4; i += blockDim.x) ones[i] = mma_fmt == FMT_F16 ? 0x3C003C00u : 0x3F803F80u;
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        const uint32_t tx = uint32_t(a_bytes + 9 * b_tile_bytes);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
            const int b = kb / P.nrb;
            const int hq0 = (kb - b * P.nrb) * P.BH;  // first padded image row of this block
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                uint8_t* sp = smem + size_t(stage) * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[stage], tx);
                tma_load_4d(sp, &tmap_x, &full_bar[stage], x_c0, 0, hq0, b);
                tma_load_4d(sp + a_atom_bytes, &tmap_x, &full_bar[stage], x_c0 + 64, 0, hq0, b);
                for (int tap = 0; tap < 9; ++tap)  // n_sign = -1: dY seen from padded position q is dY[q - (ky, kx)]
                    tma_load_4d(sp + a_bytes + tap * b_tile_bytes, &tmap_dy, &full_bar[stage], P.dy_c_off,
                                P.n_sign * (tap % 3), hq0 + P.n_sign * (tap / 3), b);
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        const uint32_t idesc = make_idesc_ab(128, n_tiles * kNarrowCo, mma_fmt, mma_fmt, 1, 1);  // both operands MN-major (K = pixels)
        const uint64_t adesc0 = make_smem_desc(smem_u32(smem), uint32_t(a_atom_bytes), 1024, kLayoutSw128);
        const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + uint32_t(a_bytes), uint32_t(b_tile_bytes), 256, kLayoutSw32);
        const uint32_t stage16 = uint32_t(stage_bytes) >> 4;
        const int ksteps = KP / 16;
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(convert ? &conv_bar[stage] : &full_bar[stage], phase);
            tc_fence_after();
            const uint64_t a_st = adesc0 + uint64_t(uint32_t(stage) * stage16);
            const uint64_t b_st = bdesc0 + uint64_t(uint32_t(stage) * stage16);
            if (elect_one()) {
                if (!(P.debug & 16)) {
                    for (int k = 0; k < ksteps; ++k)  // 16 pixels: 16 x 128 B of A, 16 x 32 B of every B tile
                        umma_f16(tmem_base, a_st + uint64_t(k * 128), b_st + uint64_t(k * 32), idesc,
                                 (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
    } else {
        // ------------------------------ epilogue: partials [tap][co][ci] ------------------------------
        if (convert) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&full_bar[stage], phase);
                uint8_t* sp = smem + size_t(stage) * stage_bytes;
                if (conv_m) convert_region_f16_to_bf16(sp, a_bytes, threadIdx.x - 64);
                else convert_region_f16_to_bf16(sp + a_bytes, 9 * b_tile_bytes, threadIdx.x - 64);  // not the ones tile
                fence_proxy_async_smem();
                mbar_arrive(&conv_bar[stage]);
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
        }
        const int q = warp & 3;
        const int m = q * 32 + lane;
        float* wsb = P.ws + (size_t(blockIdx.y) * gridDim.x + blockIdx.x) * (9 * kNarrowCo * 128);
        if (nkb > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        for (int tap = 0; tap < 9; ++tap) {
            float v[16];
            if (nkb > 0) {
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(tap * kNarrowCo), v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            float* wp = wsb + size_t(tap) * kNarrowCo * 128 + m;
#pragma unroll
            for (int i = 0; i < 16; ++i) wp[size_t(i) * 128] = v[i];
        }
        if (P.with_ones) {
            float v[16];
            v[0] = 0.f;
            if (nkb > 0) {
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(kNarrowN), v);
                tmem_ld_wait();
            }
            P.ws_bias[(size_t(blockIdx.y) * gridDim.x + blockIdx.x) * 128 + m] = v[0];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace scm
