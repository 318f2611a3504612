// Weight gradient of the 3x3 convolutions as tcgen05 GEMMs reduced over pixels (sm_100a).
//
// Replaces cuDNN's convolution_backward (wgrad half) invoked by loss.backward() (reference main.py:285)
// for every conv of models.py:51-56, 129-133, 260-266.
//
//   dW[tap][m][n] = sum_{interior pixels p} P[p (+shift)][m] * Q[p (+shift)][n]
//
// One operand is the output gradient dY (an *interior view* tensor map: everything outside the H x W
// interior is out-of-bounds and zero-filled by TMA, so halo rows never contribute), the other is the layer
// input X (full padded-plane view, box shifted by the filter tap).  Which of the two plays the M role
// (exactly 128 channels) is chosen by the host so that M = 128.  Both operands are MN-major in shared
// memory (channels contiguous, pixels along K), which is exactly what a TMA box {channels, w, h, 1}
// produces, so no transposition is ever materialised.
//
// Work split: a CTA owns a group of `tg` consecutive taps (tg*n <= 512 TMEM columns) and a contiguous
// range of pixel blocks (split-K).  The fp32 partial results are reduced with red.global.add.f32 directly
// into a gradient tensor with arbitrary (m, n, tap) strides, i.e. straight into PyTorch's
// [Cout][Cin][3][3] (or ConvTranspose [Cin][Cout][3][3]) layout.
#pragma once
#include "ptx.cuh"

namespace scm {

struct WgradParams {
    int B;
    int BW, BH;        // pixel box; KP = BW*BH pixels per K block (multiple of 16).  BW covers a full row
    int nby;           // ceil(Hd / BH), Hd = rows of P's view (H for an interior view, H+2 for a padded view)
    int num_kblocks;   // B * nby
    int kb_per_cta;    // K blocks per split
    int tap0_stride;   // taps per group (tg)
    int n;             // N (channels of Q handled by this launch), multiple of 16, <= 256, tg*n <= 512
    int q_aw;          // Q atom width in channels: 64, 32 or 16 (divides n)
    int p_c_off, q_c_off;
    int q_sign;        // Q box origin = (q_sign*kx, h0 + q_sign*ky): +1 when Q is the padded input plane and
                       // P the interior-view gradient; -1 when P is the padded input and Q the gradient
    int flip;          // 1: gradient tap index = 8 - tap (ConvTranspose / flipped packing)
    float scale;
    float* g;          // gradient tensor (fp32, accumulated atomically)
    long long g_sm, g_sn, g_st;  // element strides for m, n, tap
    int m_valid, n_valid;
    int debug;  // profiling aid (env SCMGAN_DEBUG bit3): skip the reduction
    float* ws;  // split-K partials [split][tap][n][128] (fp32) or nullptr => red.global.add straight into g
    int p_fmt, q_fmt;  // 16-bit formats of the two operand planes (FMT_BF16 / FMT_F16)
};

constexpr int kWgradThreads = 192;
constexpr int kWgradConvThreads = 128;  // warps 2..5: operand-format converters during the main loop, epilogue after it

// tcgen05.mma kind::f16 needs both operands in ONE 16-bit format (measured on B200: a bf16 A with an fp16 B raises an
// illegal-instruction fault), but the weight gradient multiplies a bf16 gradient plane by an fp16 forward activation.
// The four epilogue warps, idle during the main loop, therefore rewrite the activation tile of every pipeline stage in
// place from fp16 to bf16 (same size, elementwise, so the swizzled layout is irrelevant) between the TMA completion and
// the MMAs.  Dropping three significand bits of X only perturbs dW linearly (2^-9 per element, averaged over ~1e5
// pixels); the kink-sensitive forward pass keeps its fp16 operands.
__device__ __forceinline__ uint32_t f16x2_to_bf16x2(uint32_t w) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    const float2 f = __half22float2(h);
    return pack2_bf16(f.x, f.y);
}
__device__ __forceinline__ void convert_region_f16_to_bf16(uint8_t* base, int bytes, int tid) {
    uint4* p = reinterpret_cast<uint4*>(base);
    const int n = bytes >> 4;
    for (int i = tid; i < n; i += kWgradConvThreads) {
        uint4 v = p[i];
        v.x = f16x2_to_bf16x2(v.x); v.y = f16x2_to_bf16x2(v.y); v.z = f16x2_to_bf16x2(v.z); v.w = f16x2_to_bf16x2(v.w);
        p[i] = v;
    }
}

__global__ void __launch_bounds__(kWgradThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q,
                     const __grid_constant__ WgradParams P, int num_stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KP = P.BW * P.BH;
    const int tg = P.tap0_stride;
    const int p_atom_bytes = KP * 128;                          // one 64-channel atom of P
    const int p_bytes = 2 * p_atom_bytes;                       // M = 128
    const int q_atoms = P.n / P.q_aw;
    const int q_atom_bytes = (KP * P.q_aw * 2 + 1023) & ~1023;  // keep every atom 1024B aligned
    const int q_tap_bytes = q_atoms * q_atom_bytes;
    const int stage_bytes = p_bytes + tg * q_tap_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(num_stages) * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* acc_full = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    uint64_t* conv_bar = bars + 24;  // [8] stage converted (only used when the operand formats differ)
    // which operand (if any) has to be rewritten to the other's format: the fp16 one
    const bool conv_p = P.p_fmt != P.q_fmt && P.p_fmt == FMT_F16;
    const bool conv_q = P.p_fmt != P.q_fmt && P.q_fmt == FMT_F16;
    const bool convert = conv_p || conv_q;
    const int mma_fmt = convert ? FMT_BF16 : P.p_fmt;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int group = blockIdx.y;          // tap group
    const int tap_begin = group * tg;
    const int ntaps = min(tg, 9 - tap_begin);
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
            // warp-uniform loop, one elected lane issues the TMA loads
            const uint32_t tx = uint32_t(2 * KP * 128 + ntaps * q_atoms * KP * P.q_aw * 2);
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
                    for (int t = 0; t < ntaps; ++t) {
                        const int tap = tap_begin + t;
                        const int dy = P.q_sign * (tap / 3), dx = P.q_sign * (tap % 3);
                        uint8_t* sq = sp + p_bytes + t * q_tap_bytes;
                        for (int j = 0; j < q_atoms; ++j)
                            tma_load_4d(sq + j * q_atom_bytes, &tmap_q, &full_bar[stage], P.q_c_off + j * P.q_aw, dx,
                                        h0 + dy, b);
                    }
                }
                __syncwarp();
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
        } else if (warp == 1) {
            const uint32_t idesc = make_idesc_ab(128, P.n, mma_fmt, mma_fmt, 1, 1);
            const uint64_t q_layout = P.q_aw == 64 ? kLayoutSw128 : (P.q_aw == 32 ? kLayoutSw64 : kLayoutSw32);
            const uint32_t q_sbo = 8u * P.q_aw * 2u;
            const uint32_t q_kstep16 = (16u * P.q_aw * 2u) >> 4;
            const uint32_t q_tap16 = uint32_t(q_tap_bytes) >> 4;
            const uint64_t adesc0 = make_smem_desc(smem_u32(smem), p_atom_bytes, 1024, kLayoutSw128);
            const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + p_bytes, q_atom_bytes, q_sbo, q_layout);
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
                    for (int t = 0; t < ntaps; ++t) {
                        const uint32_t tmem_d = tmem_base + uint32_t(t * P.n);
                        const uint64_t bt = b_st + uint64_t(uint32_t(t) * q_tap16);
                        for (int k = 0; k < ksteps; ++k)
                            umma_f16(tmem_d, a_st + uint64_t(k * 128), bt + uint64_t(uint32_t(k) * q_kstep16), idesc,
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
            if (convert) {  // fp16 operand -> bf16 in place, stage by stage
                int stage = 0;
                uint32_t phase = 0;
                const int ctid = threadIdx.x - 64;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&full_bar[stage], phase);
                    uint8_t* sp = smem + size_t(stage) * stage_bytes;
                    if (conv_p) convert_region_f16_to_bf16(sp, p_bytes, ctid);
                    else convert_region_f16_to_bf16(sp + p_bytes, ntaps * q_tap_bytes, ctid);
                    fence_proxy_async_smem();
                    mbar_arrive(&conv_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
            const int q = warp & 3;
            const int m = q * 32 + lane;
            mbar_wait(acc_full, 0);
            tc_fence_after();
            for (int t = 0; t < ntaps; ++t) {
                const int tap = tap_begin + t;
                const int gtap = P.flip ? 8 - tap : tap;
                for (int n0 = 0; n0 < P.n; n0 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(t * P.n + n0), v);
                    tmem_ld_wait();
                    if (P.ws) {
                        // deterministic path: coalesced partials (m fastest), summed by wgrad_reduce_kernel
                        float* wp = P.ws + ((size_t(blockIdx.x) * 9 + tap) * P.n + n0) * 128 + m;
#pragma unroll
                        for (int i = 0; i < 16; ++i) wp[size_t(i) * 128] = v[i];
                    } else if (m < P.m_valid && !P.debug) {
                        float* gp = P.g + (long long)m * P.g_sm + (long long)gtap * P.g_st;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (n0 + i < P.n_valid) atomicAdd(gp + (long long)(n0 + i) * P.g_sn, v[i] * P.scale);
                        }
                    }
                }
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

// Sum the split-K partials and scatter into the caller's gradient layout:
//   g[m*g_sm + n*g_sn + tap'*g_st] += scale * sum_split ws[split][tap][n][m]
// Block = 32 float4 columns (512 contiguous bytes of one split row) x LANES split lanes (8, or 32 when the
// output is small and the grid would otherwise leave most SMs idle): every split lane sums the splits congruent to
// it, the partial sums are combined through shared memory in a fixed order.
// One extra block (blockIdx.x == ceil(total4 / 32)) sums the bias-gradient partials of conv_wgrad_v2.cuh when given:
//   db[m] += sum_split ws_bias[split][m]
template <int LANES>
__global__ void __launch_bounds__(32 * LANES) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int n,
                                                            float* __restrict__ g, long long g_sm, long long g_sn,
                                                            long long g_st, int flip, int m_valid, int n_valid,
                                                            float scale, const float* __restrict__ ws_bias,
                                                            float* __restrict__ db) {
    pdl_sync();
    __shared__ float4 part[LANES][32];
    constexpr int lanes = LANES;
    const int total4 = 9 * n * 32;  // float4 groups per split
    if (int(blockIdx.x) * 32 >= total4) {  // the bias block
        const int m = threadIdx.y * 32 + threadIdx.x;
        if (ws_bias == nullptr || m >= 128 || m >= m_valid) return;
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += __ldg(ws_bias + size_t(s) * 128 + m);
        db[m] += acc;
        return;
    }
    const int i4 = blockIdx.x * 32 + threadIdx.x;
    const int lane_s = threadIdx.y;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < total4) {
        const float4* src = reinterpret_cast<const float4*>(ws) + i4;
        for (int s = lane_s; s < splits; s += lanes) {
            const float4 a = __ldg(src + size_t(s) * total4);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
    }
    part[lane_s][threadIdx.x] = acc;
    __syncthreads();
    if (lane_s != 0 || i4 >= total4) return;
#pragma unroll
    for (int j = 1; j < lanes; ++j) {
        const float4 a = part[j][threadIdx.x];
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    const int m = (i4 & 31) * 4;
    const int nn = (i4 >> 5) % n;
    const int tap = i4 / (32 * n);
    if (m >= m_valid || nn >= n_valid) return;
    const int gtap = flip ? 8 - tap : tap;
    float* gp = g + (long long)m * g_sm + (long long)nn * g_sn + (long long)gtap * g_st;
    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (m + j < m_valid) gp[(long long)j * g_sm] += r[j] * scale;
}

// db[m] += scale * sum_split ws_bias[split][m]   (bias gradient partials of conv_wgrad_v2.cuh)
__global__ void wgrad_bias_reduce_kernel(const float* __restrict__ ws_bias, int splits, float* __restrict__ db,
                                         int m_valid, float scale) {
    pdl_sync();
    const int m = threadIdx.x;
    if (m >= m_valid) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += __ldg(ws_bias + size_t(s) * 128 + m);
    db[m] += acc * scale;
}

}  // namespace scm
