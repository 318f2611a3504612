// Weight gradient, second generation (layers with 128 output channels, W a multiple of 16).
//
// conv_wgrad.cuh loads the shifted input tile once per filter tap: 64 KB per 12 MMAs, which is exactly the measured
// L2->SM fabric limit (~10 TB/s chip-wide) - the kernel sat at 40 us per 128x128 layer instead of the 17 us the tensor
// core needs.  Here the input operand X is loaded once per filter ROW: one padded image row (W+2 pixels, all channels
// of the slice) per stage row, and the three kx taps address that same shared-memory tile at row offsets 0/1/2 (the
// UMMA descriptor start address advances by one 128-byte pixel row; swizzle is applied on absolute addresses, see
// conv_igemm_v2.cuh).  A stage covers BH image rows (KP = W*BH pixels) so that ~24 MMAs amortise one barrier
// round trip.  P = dY through the interior-view map (zero outside the H x W interior), Q = X through the padded view.
//
// Bias gradient for free: db[m] = sum_p dY[p][m] is one more GEMM column, dY x ones.  The CTAs of the middle filter row
// issue one extra N=16 MMA per K step against a constant all-ones tile; the kernel is bound by the L2->SM fabric, not
// by the tensor pipe, so these MMAs are hidden, and a separate pass over the 36 MB gradient plane disappears.
#pragma once
#include "conv_wgrad.cuh"

namespace scm {

struct WgradV2Params {
    int B, W;
    int BH;           // image rows per stage
    int nby;          // ceil(H / BH)
    int num_kblocks;  // B * nby
    int kb_per_cta;
    int n;            // Q channels of this launch (multiple of 16, <= 128)
    int q_aw;         // Q atom width in channels (64 / 32 / 16)
    int wq;           // rows of one Q row tile (>= W + 2, multiple of 8)
    int p_c_off, q_c_off;
    float* ws;        // split-K partials [split][tap][n][128]
    float* ws_bias;   // optional bias-gradient partials [split][128] (nullptr: not requested)
    int debug;
    int p_fmt, q_fmt;  // 16-bit formats of dY (P) and X (Q)
};

__global__ void __launch_bounds__(kWgradThreads, 1)
conv3x3_wgrad_v2_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q,
                        const __grid_constant__ WgradV2Params P, int num_stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KP = P.W * P.BH;
    const int p_atom_bytes = KP * 128;
    const int p_bytes = 2 * p_atom_bytes;
    const int q_atoms = P.n / P.q_aw;
    const int q_row_bytes = P.q_aw * 2;
    const int q_atom_bytes = (P.wq * q_row_bytes + 1023) & ~1023;
    const int q_tile_bytes = q_atoms * q_atom_bytes;          // one image row, all atoms
    const int stage_bytes = p_bytes + P.BH * q_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(num_stages) * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* acc_full = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    uint64_t* conv_bar = bars + 24;  // [8] X tile of the stage rewritten to dY's format (conv_wgrad.cuh)
    uint8_t* s_ones = reinterpret_cast<uint8_t*>(bars + 32);  // 512 B: 16 pixels x 16 channels of bf16 1.0
    const bool convert = P.p_fmt == FMT_BF16 && P.q_fmt == FMT_F16;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ky = blockIdx.y;  // this CTA's filter row: taps ky*3 + {0,1,2}
    const bool do_bias = (P.ws_bias != nullptr) && ky == 1;
    if (do_bias && threadIdx.x < 128)
        reinterpret_cast<uint32_t*>(s_ones)[threadIdx.x] = P.p_fmt == FMT_F16 ? 0x3C003C00u : 0x3F803F80u;
    if (do_bias) fence_proxy_async_smem();  // generic-proxy writes must be visible to the tensor core (async proxy)
    const int kb_begin = blockIdx.x * P.kb_per_cta;
    const int kb_end = min(P.num_kblocks, kb_begin + P.kb_per_cta);
    const int nkb = max(0, kb_end - kb_begin);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_p);
        prefetch_tmap(&tmap_q);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&conv_bar[s], kWgradConvThreads);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    pdl_sync();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nkb > 0) {
        if (warp == 0) {
            const uint32_t tx = uint32_t(2 * KP * 128 + P.BH * q_atoms * P.wq * q_row_bytes);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const int b = kb / P.nby;
                const int h0 = (kb - b * P.nby) * P.BH;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sp = smem + size_t(stage) * stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[stage], tx);
                    tma_load_4d(sp, &tmap_p, &full_bar[stage], P.p_c_off, 0, h0, b);
                    tma_load_4d(sp + p_atom_bytes, &tmap_p, &full_bar[stage], P.p_c_off + 64, 0, h0, b);
                    for (int j = 0; j < P.BH; ++j) {
                        uint8_t* sq = sp + p_bytes + j * q_tile_bytes;
                        for (int a = 0; a < q_atoms; ++a)
                            tma_load_4d(sq + a * q_atom_bytes, &tmap_q, &full_bar[stage], P.q_c_off + a * P.q_aw, 0,
                                        h0 + j + ky, b);
                    }
                }
                __syncwarp();
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
        } else if (warp == 1) {
            const uint32_t idesc = make_idesc_ab(128, P.n, P.p_fmt, convert ? FMT_BF16 : P.q_fmt, 1, 1);
            const uint64_t q_layout = P.q_aw == 64 ? kLayoutSw128 : (P.q_aw == 32 ? kLayoutSw64 : kLayoutSw32);
            const uint64_t adesc0 = make_smem_desc(smem_u32(smem), p_atom_bytes, 1024, kLayoutSw128);
            const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + p_bytes, q_atom_bytes, 8u * q_row_bytes, q_layout);
            const uint32_t stage16 = uint32_t(stage_bytes) >> 4;
            const uint32_t q_tile16 = uint32_t(q_tile_bytes) >> 4;
            const uint32_t q_row16 = uint32_t(q_row_bytes) >> 4;
            const int ksteps = P.W / 16;
            const uint32_t idesc_b = make_idesc_ab(128, 16, P.p_fmt, P.p_fmt, 1, 1);  // ones tile in dY's format
            const uint64_t ones_desc = make_smem_desc(smem_u32(s_ones), 512, 256, kLayoutSw32);
            const uint32_t tmem_b = tmem_base + uint32_t(3 * P.n);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(convert ? &conv_bar[stage] : &full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a_st = adesc0 + uint64_t(uint32_t(stage) * stage16);
                const uint64_t b_st = bdesc0 + uint64_t(uint32_t(stage) * stage16);
                if (elect_one()) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        if (P.debug & 16) break;  // profiling: TMA + barriers only
                        const uint32_t tmem_d = tmem_base + uint32_t(kx * P.n);
                        for (int j = 0; j < P.BH; ++j) {
                            const uint64_t aj = a_st + uint64_t(uint32_t(j * P.W) * 8u);            // 128 B per pixel row
                            const uint64_t bj = b_st + uint64_t(uint32_t(j) * q_tile16 + uint32_t(kx) * q_row16);
                            for (int k = 0; k < ksteps; ++k)
                                umma_f16(tmem_d, aj + uint64_t(k * 128), bj + uint64_t(uint32_t(k * 16) * q_row16), idesc,
                                         (i > 0 || j > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    if (do_bias) {
                        for (int jk = 0; jk < P.BH * ksteps; ++jk)  // pixel rows are contiguous across image rows
                            umma_f16(tmem_b, a_st + uint64_t(jk * 128), ones_desc, idesc_b, (i > 0 || jk > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full);
            __syncwarp();
        } else {
            if (convert) {
                int stage = 0;
                uint32_t phase = 0;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&full_bar[stage], phase);
                    convert_region_f16_to_bf16(smem + size_t(stage) * stage_bytes + p_bytes, P.BH * q_tile_bytes,
                                               threadIdx.x - 64);
                    fence_proxy_async_smem();
                    mbar_arrive(&conv_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
            const int q = warp & 3;
            const int m = q * 32 + lane;
            mbar_wait(acc_full, 0);
            tc_fence_after();
            for (int kx = 0; kx < 3; ++kx) {
                const int tap = ky * 3 + kx;
                for (int n0 = 0; n0 < P.n; n0 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(kx * P.n + n0), v);
                    tmem_ld_wait();
                    if (P.debug) continue;
                    float* wp = P.ws + ((size_t(blockIdx.x) * 9 + tap) * P.n + n0) * 128 + m;
#pragma unroll
                    for (int i = 0; i < 16; ++i) wp[size_t(i) * 128] = v[i];
                }
            }
            if (do_bias) {
                float v[16];
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(3 * P.n), v);
                tmem_ld_wait();
                P.ws_bias[size_t(blockIdx.x) * 128 + m] = v[0];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace scm
