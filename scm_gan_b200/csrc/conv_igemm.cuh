// 3x3 stride-1 "same" convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d / F.pad(circular) on the hot path
// (reference models.py:51-56, 76-103, 129-154, 260-283) for both the forward pass and dgrad.
//
// Data layout ("plane"): activations live in HBM as bf16 [B][H+2][W+2][Cs] (NHWC with a 1-pixel halo).
// The halo holds either zeros (zero padding) or the wrapped border (legacy circular pad-1), so the
// convolution becomes a *valid* conv over the plane.  Flattening the plane to a 2-D matrix
// [rows = B*(H+2)*(W+2)][Cs], the input row needed by output row p for filter tap (ky,kx) is simply
// p + (ky-1)*(W+2) + (kx-1): every A-operand tile is a plain 2-D TMA box at a shifted row coordinate
// (negative / past-the-end rows are zero-filled by TMA).  The kernel computes all plane rows; rows that
// fall on the halo are not stored (their accumulators are garbage by construction).
//
//   GEMM view:  D[128 rows x N] = sum_{tap, chunk} A[rows + shift(tap)][chunk] * Wt[tap][N][chunk]^T
//   M tile = 128 plane rows, N = Cout (16..256, multiple of 16), K = 9 taps x Cin (chunks of CK channels).
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = TMEM alloc + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias/activation/gate/sampling -> HBM).
// Persistent over M tiles; two TMEM accumulator stages so the epilogue of tile i overlaps tile i+1.
#pragma once
#include "ptx.cuh"

namespace scm {

enum : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_SIGMOID = 2 };

struct IgemmParams {
    // plane geometry
    int B, H, W, Hp, Wp;
    int rows;       // B*Hp*Wp
    int num_tiles;  // ceil(rows / 128)
    // GEMM
    int n;           // UMMA N (multiple of 16, <= 256); also the packed-weight rows per tap
    int cin_chunks;  // Cin_padded / CK
    int a_c_off;     // first input channel inside the A plane
    const __nv_bfloat16* a;  // the A plane itself (software-staged narrow-K path of conv_igemm_v2.cuh)
    int a_cs;                // its channel stride
    // epilogue
    float scale;               // multiplies the accumulator (1/sigma folding, loss scaling)
    const float* bias;         // [bias_n] or nullptr
    int bias_n;                // valid entries of bias (channels >= bias_n get 0)
    const float* sample_bias;  // [B][n] per-sample bias (folded action channels) or nullptr
    const float* sample_scale; // [B] per-sample factor on the accumulator (on top of `scale`) or nullptr: lets samples
                               // that belong to different spectral-norm calls (different sigma) share one GEMM
    int act;                   // ACT_*
    float slope;               // LeakyReLU negative slope
    // bf16 plane output (nullptr = none)
    __nv_bfloat16* out;
    int out_cs, out_c_off;
    int wrap;  // 1: also write the wrapped halo copies (circular); 0: write zeros on halo rows
    // optional residual add (before gating) and LeakyReLU-derivative gate, both bf16 planes (dgrad)
    const __nv_bfloat16* add;
    int add_cs, add_c_off;
    const __nv_bfloat16* gate;
    int gate_cs, gate_c_off;
    // fp32 NCHW output [B][n_valid][H][W] (nullptr = none): logits / probabilities / dz
    float* out_f32;
    int n_valid;
    // Bernoulli head (Transition conv6): z = (u < p) in training, (p > 0.5) in eval
    float* sample_out;      // fp32 NCHW [B][n_valid][H][W] or nullptr
    const float* uniforms;  // fp32 NCHW or nullptr (nullptr => threshold at 0.5)
    const unsigned long long* rng;  // device {seed, offset} for the in-kernel Philox stream, or nullptr
    int debug;              // profiling aid (env SCMGAN_DEBUG): bit0 skip global stores, bit1 skip TMEM loads
    // 16-bit formats (FMT_BF16 / FMT_F16) of the input plane, the packed weights and the output plane.  The gate plane
    // is only tested for "> 0", which reads the same bits in both formats.
    int a_fmt, b_fmt, out_fmt;
};

// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator torch/cuRAND use.  One 128-bit counter yields four
// uniforms; element idx of a tensor uses counter (offset + idx/4) and lane idx%4, so the stream does not depend on
// which thread or tile produces the element.
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long ctr, int lane) {
    uint32_t c0 = uint32_t(ctr), c1 = uint32_t(ctr >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t x = lane == 0 ? c0 : (lane == 1 ? c1 : (lane == 2 ? c2 : c3));
    return (float(x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // 24 random bits, strictly inside (0, 1)
}

// Bernoulli / threshold head shared by all conv kernels: training with injected uniforms, training with the
// in-kernel Philox stream, or evaluation (p > 0.5).
__device__ __forceinline__ float bernoulli_head(const IgemmParams& P, size_t idx, float p) {
    if (P.uniforms) return __ldg(P.uniforms + idx) < p ? 1.f : 0.f;
    if (P.rng) return philox_uniform(__ldg(P.rng), __ldg(P.rng + 1) + (idx >> 2), int(idx & 3)) < p ? 1.f : 0.f;
    return p > 0.5f ? 1.f : 0.f;
}

template <int CK>
struct IgemmCfg {
    static constexpr int kRowBytes = CK * 2;
    static constexpr uint64_t kLayout = (CK == 64) ? kLayoutSw128 : (CK == 32 ? kLayoutSw64 : kLayoutSw32);
    static constexpr int kSbo = 8 * kRowBytes;  // 8-row core-matrix group stride
    static constexpr int kATileBytes = 128 * kRowBytes;
    static constexpr int kKSteps = CK / 16;  // UMMA_K = 16 for 16-bit operands
};

__host__ __device__ inline int igemm_b_tile_bytes(int n, int ck) { return ((n * ck * 2) + 1023) & ~1023; }

constexpr int kIgemmThreads = 192;
constexpr int kAccStageCols = 256;
constexpr int kMaxStages = 8;

template <int CK>
__global__ void __launch_bounds__(kIgemmThreads, 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ IgemmParams P, int num_stages) {
    using Cfg = IgemmCfg<CK>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A | B)] then barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_tile_bytes = igemm_b_tile_bytes(P.n, CK);
    const int stage_bytes = Cfg::kATileBytes + b_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(num_stages) * stage_bytes);
    uint64_t* full_bar = bars;                     // [num_stages]
    uint64_t* empty_bar = bars + kMaxStages;       // [num_stages]
    uint64_t* acc_full = bars + 2 * kMaxStages;    // [2]
    uint64_t* acc_empty = bars + 2 * kMaxStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    pdl_sync();  // everything above is CTA-local: it overlaps the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int k_iters = 9 * P.cin_chunks;
    const uint32_t tx_bytes = uint32_t(Cfg::kATileBytes + P.n * Cfg::kRowBytes);

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
                const int m0 = tile * 128;
                for (int tap = 0; tap < 9; ++tap) {
                    const int shift = (tap / 3 - 1) * P.Wp + (tap % 3 - 1);
                    for (int c = 0; c < P.cin_chunks; ++c) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + size_t(stage) * stage_bytes;
                        uint8_t* sb = sa + Cfg::kATileBytes;
                        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], P.a_c_off + c * CK, m0 + shift);
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], c * CK, tap * P.n);
                        if (++stage == num_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            const uint32_t idesc = make_idesc_ab(128, P.n, P.a_fmt, P.b_fmt, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * kAccStageCols);
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + size_t(stage) * stage_bytes);
                    const uint32_t sb = sa + Cfg::kATileBytes;
                    const uint64_t adesc = make_smem_desc(sa, 16, Cfg::kSbo, Cfg::kLayout);
                    const uint64_t bdesc = make_smem_desc(sb, 16, Cfg::kSbo, Cfg::kLayout);
#pragma unroll
                    for (int k = 0; k < Cfg::kKSteps; ++k) {
                        // advance 16 elements (32 bytes) along K inside the swizzle row
                        umma_f16(tmem_d, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc,
                                 (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        const int plane = P.Hp * P.Wp;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            const int p = tile * 128 + q * 32 + lane;
            const bool valid = p < P.rows;
            int b = 0, hp = 0, wp = 0;
            if (valid) {
                b = p / plane;
                const int rem = p - b * plane;
                hp = rem / P.Wp;
                wp = rem - hp * P.Wp;
            }
            const bool interior = valid && hp >= 1 && hp <= P.H && wp >= 1 && wp <= P.W;

            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kAccStageCols);

            // wrapped destinations for circular planes
            int hp2 = -1, wp2 = -1;
            if (P.wrap && interior) {
                if (hp == 1) hp2 = P.H + 1; else if (hp == P.H) hp2 = 0;
                if (wp == 1) wp2 = P.W + 1; else if (wp == P.W) wp2 = 0;
            }

            for (int n0 = 0; n0 < P.n; n0 += 16) {
                float v[16];
                if (!(P.debug & 2)) {
                    tmem_ld16(taddr + uint32_t(n0), v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = float(i);
                }
                if (!valid || (P.debug & 1)) continue;
                const float rs = P.sample_scale ? P.scale * __ldg(P.sample_scale + b) : P.scale;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = v[i] * rs;
                    if (P.bias && n0 + i < P.bias_n) x += __ldg(P.bias + n0 + i);
                    if (P.sample_bias) x += __ldg(P.sample_bias + size_t(b) * P.n + n0 + i);
                    v[i] = x;
                }
                if (interior && P.add) {
                    const uint4* ap =
                        reinterpret_cast<const uint4*>(P.add + size_t(p) * P.add_cs + P.add_c_off + n0);
                    uint4 r0 = __ldg(ap), r1 = __ldg(ap + 1);
                    const __nv_bfloat16* h0 = reinterpret_cast<const __nv_bfloat16*>(&r0);
                    const __nv_bfloat16* h1 = reinterpret_cast<const __nv_bfloat16*>(&r1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] += __bfloat162float(h0[i]);
                        v[8 + i] += __bfloat162float(h1[i]);
                    }
                }
                if (P.act == ACT_LRELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * P.slope;
                } else if (P.act == ACT_SIGMOID) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 1.f / (1.f + __expf(-v[i]));
                }
                if (interior && P.gate) {
                    const uint4* gp =
                        reinterpret_cast<const uint4*>(P.gate + size_t(p) * P.gate_cs + P.gate_c_off + n0);
                    uint4 r0 = __ldg(gp), r1 = __ldg(gp + 1);
                    const __nv_bfloat16* h0 = reinterpret_cast<const __nv_bfloat16*>(&r0);
                    const __nv_bfloat16* h1 = reinterpret_cast<const __nv_bfloat16*>(&r1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] *= (__bfloat162float(h0[i]) > 0.f) ? 1.f : P.slope;
                        v[8 + i] *= (__bfloat162float(h1[i]) > 0.f) ? 1.f : P.slope;
                    }
                }
                if (P.out) {
                    uint4 o0, o1;
                    uint32_t* w0 = reinterpret_cast<uint32_t*>(&o0);
                    uint32_t* w1 = reinterpret_cast<uint32_t*>(&o1);
                    if (interior) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            w0[i] = pack2_fmt(v[2 * i], v[2 * i + 1], P.out_fmt);
                            w1[i] = pack2_fmt(v[8 + 2 * i], v[8 + 2 * i + 1], P.out_fmt);
                        }
                    } else {
                        o0 = make_uint4(0, 0, 0, 0);
                        o1 = o0;
                    }
                    if (interior || !P.wrap) {
                        uint4* op = reinterpret_cast<uint4*>(P.out + size_t(p) * P.out_cs + P.out_c_off + n0);
                        op[0] = o0;
                        op[1] = o1;
                    }
                    if (hp2 >= 0) {
                        const size_t r = size_t(b) * plane + size_t(hp2) * P.Wp + wp;
                        uint4* op = reinterpret_cast<uint4*>(P.out + r * P.out_cs + P.out_c_off + n0);
                        op[0] = o0;
                        op[1] = o1;
                    }
                    if (wp2 >= 0) {
                        const size_t r = size_t(b) * plane + size_t(hp) * P.Wp + wp2;
                        uint4* op = reinterpret_cast<uint4*>(P.out + r * P.out_cs + P.out_c_off + n0);
                        op[0] = o0;
                        op[1] = o1;
                    }
                    if (hp2 >= 0 && wp2 >= 0) {
                        const size_t r = size_t(b) * plane + size_t(hp2) * P.Wp + wp2;
                        uint4* op = reinterpret_cast<uint4*>(P.out + r * P.out_cs + P.out_c_off + n0);
                        op[0] = o0;
                        op[1] = o1;
                    }
                }
                if (P.out_f32 && interior) {
                    const size_t hw = size_t(P.H) * P.W;
                    const size_t base = (size_t(b) * P.n_valid) * hw + size_t(hp - 1) * P.W + (wp - 1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = n0 + i;
                        if (n < P.n_valid) {
                            const size_t idx = base + size_t(n) * hw;
                            P.out_f32[idx] = v[i];
                            if (P.sample_out) P.sample_out[idx] = bernoulli_head(P, idx, v[i]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
