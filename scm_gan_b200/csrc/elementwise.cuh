// HBM-bound helper kernels around the tcgen05 GEMMs: layout packing, weight packing, spectral-norm power
// iteration and its backward, bias/colsum reductions, fused losses and the fused clip+Adam optimiser.
#pragma once
#include "ptx.cuh"

namespace scm {

// ----------------------------------------------------------------------------------------------
// fp32 NCHW (caller tensors) -> bf16 plane [B][H+2][W+2][Cs] at channel offset c_off, halo = wrap or zero.
// One thread per plane pixel; channels [c_off, c_off + c_pad) are written (zeros beyond C).
// ----------------------------------------------------------------------------------------------
// Optional `sig` (fp32 NCHW, dense): multiply by sig*(1-sig), i.e. the sigmoid derivative of a saved
// probability map (backward of the Encoder / Transition output sigmoid, reference models.py:103,154).
__global__ void pack_nchw_to_plane_kernel(const float* __restrict__ src, long long src_bstride, int C, int B, int H,
                                          int W, __nv_bfloat16* __restrict__ dst, int Cs, int c_off, int c_pad,
                                          int wrap, const float* __restrict__ sig, int fmt) {
    pdl_sync();
    const int Hp = H + 2, Wp = W + 2;
    const long long rows = (long long)B * Hp * Wp;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows) return;
    const int b = int(p / (Hp * Wp));
    const int rem = int(p - (long long)b * Hp * Wp);
    int hp = rem / Wp, wp = rem - hp * Wp;
    int h = hp - 1, w = wp - 1;
    bool zero = false;
    if (wrap) {
        h = (h + H) % H;
        w = (w + W) % W;
    } else {
        zero = (h < 0 || h >= H || w < 0 || w >= W);
    }
    const float* s = src + (long long)b * src_bstride + (long long)h * W + w;
    const float* sg = sig ? sig + ((long long)b * C * H + h) * W + w : nullptr;
    __nv_bfloat16* d = dst + p * Cs + c_off;
    for (int c0 = 0; c0 < c_pad; c0 += 8) {
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
        float xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            float x = 0.f;
            if (!zero && c < C) {
                x = __ldg(s + (long long)c * H * W);
                if (sg) {
                    const float q = __ldg(sg + (long long)c * H * W);
                    x *= q * (1.f - q);
                }
            }
            xv[i] = x;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) ow[i] = pack2_fmt(xv[2 * i], xv[2 * i + 1], fmt);
        *reinterpret_cast<uint4*>(d + c0) = o;
    }
}

// CoordConv coordinate channels: x coordinate -1 + 2w/W at channel c_off, y coordinate -1 + 2h/H at c_off + 1 of every
// interior pixel (reference coordconv.py:10-14).
__global__ void pack_coords_kernel(__nv_bfloat16* __restrict__ dst, int Cs, int c_off, int B, int H, int W, int fmt) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * H * W) return;
    const int w = int(i % W), h = int((i / W) % H), b = int(i / ((long long)W * H));
    const long long p = ((long long)b * (H + 2) + (h + 1)) * (W + 2) + (w + 1);
    *reinterpret_cast<uint32_t*>(dst + p * Cs + c_off) =
        pack2_fmt(-1.f + 2.f * float(w) / float(W), -1.f + 2.f * float(h) / float(H), fmt);
}

// ----------------------------------------------------------------------------------------------
// Weight packing: fp32 parameter tensor -> bf16 [9][n_pad][k_pad] (K-major B operand of the implicit GEMM).
//   out[tap][n][k] = W[n*s_n + k*s_k + (flip ? 8-tap : tap)] / sigma      (zero outside n_valid x k_valid)
// Covers Conv2d forward (n=co,k=ci), Conv2d dgrad (n=ci,k=co,flip), ConvTranspose2d forward/dgrad.
// ----------------------------------------------------------------------------------------------
struct PackJob {
    const float* w;
    __nv_bfloat16* out;
    const float* sigma;  // nullptr => 1
    int n_pad, k_pad, n_valid, k_valid;
    long long s_n, s_k;
    int k_src_off;  // first source k (e.g. skip nothing: 0)
    int flip;
    int out_ld;     // row pitch of out in elements (>= k_pad)
    int fmt;        // FMT_BF16 / FMT_F16
};
constexpr int kMaxPackJobs = 16;
struct PackJobs {
    PackJob job[kMaxPackJobs];
    int count;
};

__global__ void pack_weights_kernel(const __grid_constant__ PackJobs jobs) {
    pdl_sync();
    const PackJob& J = jobs.job[blockIdx.y];
    const long long total = 9LL * J.n_pad * J.k_pad;
    const float inv = J.sigma ? 1.f / __ldg(J.sigma) : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = int(i % J.k_pad);
        const long long r = i / J.k_pad;
        const int n = int(r % J.n_pad);
        const int tap = int(r / J.n_pad);
        float x = 0.f;
        if (n < J.n_valid && k < J.k_valid)
            x = __ldg(J.w + n * J.s_n + (long long)(k + J.k_src_off) * J.s_k + (J.flip ? 8 - tap : tap)) * inv;
        if (J.fmt == FMT_F16)
            reinterpret_cast<__half*>(J.out)[r * J.out_ld + k] = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
        else
            J.out[r * J.out_ld + k] = __float2bfloat16_rn(x);
    }
}

// ----------------------------------------------------------------------------------------------
// Spectral norm power iteration (reference spectral_normalization.py:23-35), one CTA per wrapped conv:
//   v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v)        (eps = 1e-12, as l2normalize)
// W is [rows][cols] fp32 (rows = Cout, cols = Cin*9).  u, v are updated in place; u, v, sigma are also
// copied to the per-call save slots needed by the backward pass.
// ----------------------------------------------------------------------------------------------
struct SnLayer {
    const float* w;
    float* u;
    float* v;
    float* sigma;   // [1]
    float* u_save;  // [rows] or nullptr
    float* v_save;  // [cols] or nullptr
    int rows, cols;
};
constexpr int kMaxSnLayers = 8;
struct SnLayers {
    SnLayer layer[kMaxSnLayers];
    int count;
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// block-wide sum for 1024 threads; result broadcast to all threads
__device__ __forceinline__ float block_sum(float x, float* red /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    if (warp == 0) {
        float y = (lane < (blockDim.x >> 5)) ? red[lane] : 0.f;
        y = warp_sum(y);
        if (lane == 0) red[32] = y;
    }
    __syncthreads();
    return red[32];
}

__global__ void __launch_bounds__(1024, 1) sn_power_iter_kernel(const __grid_constant__ SnLayers L) {
    pdl_sync();
    extern __shared__ float sn_smem[];  // [cols] t/v  + [rows] s/u + 33
    const SnLayer& Y = L.layer[blockIdx.x];
    const int R = Y.rows, C = Y.cols;
    float* sv = sn_smem;
    float* su = sn_smem + C;
    float* red = su + R;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int n = tid; n < R; n += nt) su[n] = Y.u[n];
    __syncthreads();
    // t = W^T u  (thread per column: coalesced along k)
    float ss = 0.f;
    for (int k = tid; k < C; k += nt) {
        float acc = 0.f;
        for (int n = 0; n < R; ++n) acc = fmaf(__ldg(Y.w + (long long)n * C + k), su[n], acc);
        sv[k] = acc;
        ss += acc * acc;
    }
    float nrm = sqrtf(block_sum(ss, red));
    const float inv_v = 1.f / (nrm + 1e-12f);
    for (int k = tid; k < C; k += nt) sv[k] *= inv_v;
    __syncthreads();
    // s = W v  (warp per row: coalesced along k)
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int n = warp; n < R; n += nw) {
        float acc = 0.f;
        for (int k = lane; k < C; k += 32) acc = fmaf(__ldg(Y.w + (long long)n * C + k), sv[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) su[n] = acc;
    }
    __syncthreads();
    float s2 = 0.f;
    for (int n = tid; n < R; n += nt) s2 += su[n] * su[n];
    s2 = block_sum(s2, red);
    const float inv_u = 1.f / (sqrtf(s2) + 1e-12f);
    // sigma = u_new . (W v) = sum s^2 * inv_u
    if (tid == 0) *Y.sigma = s2 * inv_u;
    for (int n = tid; n < R; n += nt) {
        const float un = su[n] * inv_u;
        Y.u[n] = un;
        if (Y.u_save) Y.u_save[n] = un;
    }
    for (int k = tid; k < C; k += nt) {
        Y.v[k] = sv[k];
        if (Y.v_save) Y.v_save[k] = sv[k];
    }
}

// Cluster version of the power iteration: eight CTAs (one thread-block cluster) share one layer.  CTA r owns the
// column slice [r*cpc, (r+1)*cpc) of W: it computes its slice of t = W^T u, its partial of |t|^2 and its partial of
// s = W v; the partials are exchanged through distributed shared memory (every CTA stores its partial into all
// peers) and summed in rank order, so the result does not depend on scheduling.  W is read by eight SMs instead of
// one (a single CTA is latency-bound: measured 45 us per call at 128 x 2304), twice, the second time from L1/L2.
constexpr int kSnCluster = 8;
struct SnClusterGeom {
    int cpc[kMaxSnLayers];     // columns per CTA
    int groups[kMaxSnLayers];  // row groups of phase 1 (threads = groups x padded slice)
    int cpc_max, rows_max;
};

// `iters` successive power iterations in ONE launch (sigma of iteration i goes to sigma[i * sigma_stride]): the power
// iteration depends only on the weights, which are constant within a training iteration, so all the per-call
// iterations of one rollout (reference: one per SpectralNorm forward call, 5T+3 per iteration) can be run ahead.
__global__ void __cluster_dims__(kSnCluster, 1, 1) __launch_bounds__(1024, 1)
sn_power_iter_cluster_kernel(const __grid_constant__ SnLayers L, const __grid_constant__ SnClusterGeom Gm, int iters,
                             int sigma_stride) {
    pdl_sync();
    extern __shared__ float sn_smem[];
    const int li = blockIdx.x / kSnCluster;
    const uint32_t rank = cluster_ctarank();
    const SnLayer& Y = L.layer[li];
    const int R = Y.rows, C = Y.cols;
    const int cpc = Gm.cpc[li], G = Gm.groups[li];
    const int cpcp = (cpc + 31) & ~31;
    const int k0 = int(rank) * cpc;
    const int nk = max(0, min(C, k0 + cpc) - k0);
    // shared layout (sized for the largest layer of the launch; identical in every CTA so peers can address it)
    float* su = sn_smem;                              // [rows_max]  u, later s
    float* st = su + Gm.rows_max;                     // [cpc_max]   t / v slice
    float* sp = st + Gm.cpc_max;                      // [rows_max]  this CTA's partial of s
    float* s_part = sp + Gm.rows_max;                 // [8][rows_max] partials of s from every rank
    float* ss_part = s_part + kSnCluster * Gm.rows_max;  // [8]
    float* red = ss_part + kSnCluster;                // [33]
    float* tpart = red + 33;                          // [groups][cpc_max]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int n = tid; n < R; n += nt) su[n] = Y.u[n];
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        // phase 1: t = W^T u on the slice; thread (g, kk) sums the rows of group g for column k0 + kk
        {
            const int kk = tid % cpcp, g = tid / cpcp;
            if (g < G && kk < nk) {
                const int rpg = (R + G - 1) / G;
                const int n1 = min(R, (g + 1) * rpg);
                const float* wp = Y.w + k0 + kk;
                float acc = 0.f;
#pragma unroll 8
                for (int n = g * rpg; n < n1; ++n) acc = fmaf(__ldg(wp + (long long)n * C), su[n], acc);
                tpart[g * Gm.cpc_max + kk] = acc;
            }
        }
        __syncthreads();
        float ss = 0.f;
        if (tid < nk) {
            float t = 0.f;
            for (int g = 0; g < G; ++g) t += tpart[g * Gm.cpc_max + tid];
            st[tid] = t;
            ss = t * t;
        }
        ss = block_sum(ss, red);
        if (tid < kSnCluster) {
            const uint32_t remote = mapa_shared(smem_u32(ss_part + rank), uint32_t(tid));
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(ss) : "memory");
        }
        cluster_sync_all();
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < kSnCluster; ++r) tot += ss_part[r];
        const float inv_v = 1.f / (sqrtf(tot) + 1e-12f);
        if (tid < nk) st[tid] *= inv_v;
        __syncthreads();
        // phase 2: partial s = W[:, slice] v[slice]  (warp per row)
        for (int n = warp; n < R; n += nw) {
            const float* wp = Y.w + (long long)n * C + k0;
            float acc = 0.f;
            for (int kk = lane; kk < nk; kk += 32) acc = fmaf(__ldg(wp + kk), st[kk], acc);
            acc = warp_sum(acc);
            if (lane == 0) sp[n] = acc;
        }
        __syncthreads();
        for (int i = tid; i < kSnCluster * R; i += nt) {
            const int peer = i / R, n = i - peer * R;
            const uint32_t remote = mapa_shared(smem_u32(s_part + int(rank) * Gm.rows_max + n), uint32_t(peer));
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(sp[n]) : "memory");
        }
        cluster_sync_all();
        float s2 = 0.f;
        for (int n = tid; n < R; n += nt) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < kSnCluster; ++r) s += s_part[r * Gm.rows_max + n];
            su[n] = s;
            s2 += s * s;
        }
        s2 = block_sum(s2, red);
        const float inv_u = 1.f / (sqrtf(s2) + 1e-12f);
        if (rank == 0 && tid == 0) Y.sigma[(long long)it * sigma_stride] = s2 * inv_u;  // sigma = u_new . (W v)
        for (int n = tid; n < R; n += nt) su[n] *= inv_u;  // every CTA holds the full new u for the next iteration
        if (it + 1 < iters) cluster_sync_all();  // peers are done reading ss_part / s_part before they are rewritten
        else __syncthreads();
    }
    if (rank == 0) {
        for (int n = tid; n < R; n += nt) {
            Y.u[n] = su[n];
            if (Y.u_save) Y.u_save[n] = su[n];
        }
    }
    if (tid < nk) {
        Y.v[k0 + tid] = st[tid];
        if (Y.v_save) Y.v_save[k0 + tid] = st[tid];
    }
}

// Spectral-norm backward: dWbar = G/sigma - (<G, Wbar>/sigma^2) * u v^T   (u, v, sigma of that forward call)
struct SnBwdLayer {
    const float* g;     // gradient wrt the normalised weight [rows][cols]
    const float* wbar;  // [rows][cols]
    const float* u;
    const float* v;
    const float* sigma;
    float* dot;   // [1] scratch, zeroed by the caller
    float* out;   // [rows][cols]
    int rows, cols;
    int accumulate;  // out += result instead of out = result
    const float* sigma2;  // optional (see scmgan_sn_bwd_layer): g was taken with respect to Wbar/sigma, sigma2 is the
                          // sigma of the call the samples belong to; nullptr = sigma
};
struct SnBwdLayers {
    SnBwdLayer layer[kMaxSnLayers];
    int count;
};

__global__ void sn_bwd_dot_kernel(const __grid_constant__ SnBwdLayers L) {
    pdl_sync();
    __shared__ float red[33];
    const SnBwdLayer& Y = L.layer[blockIdx.y];
    const long long total = (long long)Y.rows * Y.cols;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        acc = fmaf(__ldg(Y.g + i), __ldg(Y.wbar + i), acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(Y.dot, acc);
}

__global__ void sn_bwd_apply_kernel(const __grid_constant__ SnBwdLayers L) {
    pdl_sync();
    const SnBwdLayer& Y = L.layer[blockIdx.y];
    const long long total = (long long)Y.rows * Y.cols;
    const float sig = __ldg(Y.sigma);
    const float inv = 1.f / sig;
    const float coef = __ldg(Y.dot) * inv / (Y.sigma2 ? __ldg(Y.sigma2) : sig);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n = int(i / Y.cols), k = int(i - (long long)n * Y.cols);
        const float r = __ldg(Y.g + i) * inv - coef * __ldg(Y.u + n) * __ldg(Y.v + k);
        Y.out[i] = Y.accumulate ? Y.out[i] + r : r;
    }
}

// ----------------------------------------------------------------------------------------------
// Per-sample channel sums over the interior of a bf16 plane (bias gradients / folded-action gradients):
//   S[b][c] += sum_{interior p} plane[b][p][c_off + c];   db[c] += same summed over b (optional)
// grid = (row chunks, B); block = 256 threads = (n/8 channel groups) x (row lanes)
// ----------------------------------------------------------------------------------------------
__global__ void plane_colsum_kernel(const __nv_bfloat16* __restrict__ plane, int Cs, int c_off, int n, int B, int H,
                                    int W, float* __restrict__ S, float* __restrict__ db, int rows_per_block) {
    pdl_sync();
    extern __shared__ float cs_smem[];  // [row lanes][n]
    const int Hp = H + 2, Wp = W + 2;
    const int groups = n >> 3;
    const int lanes = blockDim.x / groups;
    const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
    const int b = blockIdx.y;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rl < lanes) {
        const int r0 = blockIdx.x * rows_per_block;
        const int r1 = min(H * W, r0 + rows_per_block);
        for (int r = r0 + rl; r < r1; r += lanes) {
            const int h = r / W, w = r - h * W;
            const long long p = ((long long)b * Hp + (h + 1)) * Wp + (w + 1);
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(plane + p * Cs + c_off + g * 8));
            const __nv_bfloat16* hq = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += __bfloat162float(hq[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cs_smem[rl * n + g * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += cs_smem[l * n + c];
        if (S) atomicAdd(S + (long long)b * n + c, s);
        if (db) atomicAdd(db + c, s);
    }
}

// ----------------------------------------------------------------------------------------------
// Folded action channels of Transition.conv1 (reference models.py:69-73): under circular padding a
// spatially constant input channel contributes a per-sample constant, so
//   sample_bias[b][n] = bias[n] + (1/sigma) * sum_a act[b][a] * sum_tap Wbar[n][L + a][tap]
// and in backward   dW[n][L + a][tap] = sum_b S[b][n] * act[b][a]   for every tap.
// ----------------------------------------------------------------------------------------------
__global__ void action_bias_kernel(const float* __restrict__ wbar, const float* __restrict__ sigma,
                                   const float* __restrict__ bias, const float* __restrict__ act, int B, int Cout,
                                   int L, int A, float* __restrict__ out) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Cout) return;
    const int b = i / Cout, n = i - b * Cout;
    const float inv = sigma ? 1.f / __ldg(sigma) : 1.f;
    float acc = 0.f;
    for (int a = 0; a < A; ++a) {
        const float av = __ldg(act + b * A + a);
        if (av != 0.f) {
            const float* wp = wbar + ((long long)n * (L + A) + (L + a)) * 9;
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) s += __ldg(wp + t);
            acc = fmaf(av, s, acc);
        }
    }
    out[i] = (bias ? __ldg(bias + n) : 0.f) + acc * inv;
}

__global__ void action_wgrad_kernel(const float* __restrict__ S, const float* __restrict__ act, int B, int Cout,
                                    int L, int A, float* __restrict__ g) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * A) return;
    const int n = i / A, a = i - n * A;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(__ldg(S + b * Cout + n), __ldg(act + b * A + a), acc);
    float* gp = g + ((long long)n * (L + A) + (L + a)) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) gp[t] = acc;
}

// ----------------------------------------------------------------------------------------------
// Fused sigmoid + binary cross entropy + masked per-sample mean (reference main.py:188-197, 310-312):
//   loss += (1/B) * sum_b mask[b] * mean_{c,h,w} BCE(sigmoid(x), y)       (log clamped at -100 as torch)
//   dx = (sigmoid(x) - y) * mask[b] / (B*C*H*W)
// The gradient is produced in the same pass (forward+backward fusion); autograd scales it by grad_output.
// ----------------------------------------------------------------------------------------------
// Sequence form: x holds the logits of T rollout steps, t-major ([T][B][per]); step t's target rows are
// y + b*y_bstride + t*y_tstride (a [B, T] window of the [B][Hn][C][H][W] frame tensor) and its mask entries
// mask + b*m_bstride + t*m_tstride; loss_t[t] += the step's term (T = 1: the single-step op).  The decoder has no
// state, so the T decodes of one iteration (main.py:188-197) run as ONE batch of T*B samples.
__global__ void bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, long long y_bstride,
                                  long long y_tstride, const float* __restrict__ mask, long long m_bstride,
                                  long long m_tstride, int T, int B, long long per, float* __restrict__ loss_t,
                                  float* __restrict__ dx) {
    pdl_sync();
    __shared__ float red[33];
    const int row = blockIdx.y;  // t * B + b
    const int t = row / B, b = row - t * B;
    const float m = mask ? __ldg(mask + (long long)b * m_bstride + (long long)t * m_tstride) : 1.f;
    const float inv = 1.f / (float(B) * float(per));
    const float* xb = x + (long long)row * per;
    const float* yb = y + (long long)b * y_bstride + (long long)t * y_tstride;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const float xv = __ldg(xb + i), yv = __ldg(yb + i);
        const float p = 1.f / (1.f + __expf(-xv));
        const float lp = fmaxf(__logf(p), -100.f), lq = fmaxf(__logf(1.f - p), -100.f);
        acc -= yv * lp + (1.f - yv) * lq;
        if (dx) dx[(long long)row * per + i] = (p - yv) * m * inv;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss_t + t, acc * m * inv);
}

// ----------------------------------------------------------------------------------------------
// Fused decoder loss head, second half: loss_t[t] += sum over the per-warp partial rows written by the conv epilogue
// (igemm_epilogue.cuh, BceCarry), in a fixed order.
// ----------------------------------------------------------------------------------------------
__global__ void bce_finalize_kernel(const float* __restrict__ ws, int rows, int T, float* __restrict__ loss_t) {
    pdl_sync();
    __shared__ float red[33];
    const int t = blockIdx.x;   // one block per rollout step
    float acc = 0.f;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) acc += ws[(long long)r * T + t];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) loss_t[t] += acc;
}

// ----------------------------------------------------------------------------------------------
// Backward companion of the fused decoder loss head: the plane holds d loss_t / d logits of every rollout step t
// (t-major); the chain rule needs g[t] = d total / d loss_t on top.  The training step sums the terms (g = 1), so the
// kernel reads g[t] first and leaves step t alone when it is exactly 1 - one 4-byte load per block instead of a
// read-modify-write pass over the plane.
// ----------------------------------------------------------------------------------------------
__global__ void plane_scale_steps_kernel(uint4* __restrict__ plane, long long vec_per_step, const float* __restrict__ g,
                                         int fmt) {
    pdl_sync();
    const float k = __ldg(g + blockIdx.y);
    if (k == 1.f) return;
    uint4* base = plane + (long long)blockIdx.y * vec_per_step;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_step;
         i += (long long)gridDim.x * blockDim.x) {
        uint4 q = base[i];
        uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float lo, hi;
            if (fmt == FMT_F16) {
                const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
                lo = __low2float(h); hi = __high2float(h);
            } else {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                lo = __low2float(h); hi = __high2float(h);
            }
            w[j] = pack2_fmt(lo * k, hi * k, fmt);
        }
        base[i] = q;
    }
}

// ----------------------------------------------------------------------------------------------
// Weight gradient of the two CoordConv coordinate channels.  The coordinates are generated inside the convolution's
// im2col tile and never stored, so the tensor-core weight-gradient kernels (which read the input plane) see zeros there;
// their gradient is a plain correlation of dy with two known ramps:
//   g[co][coord_c + c][ky][kx] = sum_{b,h,w} dy[b][co][h][w] * coord_c(h + ky - 1, w + kx - 1)   (0 outside the image)
// dy: fp32 NCHW (the op's incoming gradient).  One block per output channel, fixed-order reduction (deterministic).
__global__ void coord_wgrad_kernel(const float* __restrict__ dy, int B, int Co, int H, int W, float* __restrict__ g,
                                   long long g_s_co, long long g_s_ci, int coord_c) {
    pdl_sync();
    __shared__ float red[33];
    const int co = blockIdx.x;
    float acc[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) acc[i] = 0.f;
    const long long n = (long long)B * H * W;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const int w = int(i % W), h = int((i / W) % H), b = int(i / ((long long)W * H));
        const float d = __ldg(dy + (((long long)b * Co + co) * H + h) * W + w);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int hh = h + ky - 1, ww = w + kx - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                    acc[ky * 3 + kx] = fmaf(d, -1.f + 2.f * float(ww) / float(W), acc[ky * 3 + kx]);
                    acc[9 + ky * 3 + kx] = fmaf(d, -1.f + 2.f * float(hh) / float(H), acc[9 + ky * 3 + kx]);
                }
            }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        const float s = block_sum(acc[i], red);
        if (threadIdx.x == 0) g[(long long)co * g_s_co + (long long)(coord_c + i / 9) * g_s_ci + (i % 9)] = s;
    }
}

// ----------------------------------------------------------------------------------------------
// Rollout-MSE evaluation (reference measure_prediction_mse, main.py:784-836).  The decoder / reward predictor are
// stateless, so all steps of the evaluation rollout are decoded as one batch; these two kernels turn the logits into the
// four per-step curves without a single host round trip (the reference reads four scalars back per step).
//   eval_sqerr_kernel: out[t*B + b] = mean_chw (y[b,t] - sigmoid(x[t,b]))^2     (main.py:812-815, before the mask)
//   eval_stats_kernel: one block walks the steps in order; mask_t = prod_{s<=t} (1 - done_s) (main.py:806),
//       d = mask * sqerr, r = mask * (sum_r reward - sum_r predicted)^2 (main.py:822-824),
//       table[t] = { mean(d) B/live, std(d) B/live, mean(r) B/live, std(r) B/live, live }   (std unbiased: torch.std)
// One block per row / one block in total: fixed summation order, deterministic.
// ----------------------------------------------------------------------------------------------
__global__ void eval_sqerr_kernel(const float* __restrict__ x, const float* __restrict__ y, long long y_bstride,
                                  long long y_tstride, int B, long long per, float* __restrict__ out) {
    pdl_sync();
    __shared__ float red[33];
    const long long row = blockIdx.x;  // t * B + b
    const int t = int(row / B), b = int(row - (long long)t * B);
    const float* xb = x + row * per;
    const float* yb = y + (long long)b * y_bstride + (long long)t * y_tstride;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) {
        const float d = __ldg(yb + i) - 1.f / (1.f + __expf(-__ldg(xb + i)));
        acc = fmaf(d, d, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[row] = acc / float(per);
}

__global__ void eval_stats_kernel(const float* __restrict__ sqerr, const float* __restrict__ rpred,
                                  const float* __restrict__ rewards, long long r_bstride, long long r_tstride,
                                  const float* __restrict__ dones, long long d_bstride, long long d_tstride, int T,
                                  int B, int R, float* __restrict__ table) {
    pdl_sync();
    extern __shared__ float ev_smem[];  // [B] running mask, [B] d, [B] r, 33 reduction slots
    float* mask = ev_smem;
    float* dv = mask + B;
    float* rv = dv + B;
    float* red = rv + B;
    for (int b = threadIdx.x; b < B; b += blockDim.x) mask[b] = 1.f;
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        float live = 0.f, sd = 0.f, sr = 0.f;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float m = mask[b] * (1.f - __ldg(dones + (long long)b * d_bstride + (long long)t * d_tstride));
            mask[b] = m;
            float re = 0.f, rp = 0.f;
            for (int r = 0; r < R; ++r) {
                re += __ldg(rewards + (long long)b * r_bstride + (long long)t * r_tstride + r);
                rp += rpred[((long long)t * B + b) * R + r];
            }
            const float d = m * sqerr[(long long)t * B + b];
            const float q = m * (re - rp) * (re - rp);
            dv[b] = d; rv[b] = q;
            live += m; sd += d; sr += q;
        }
        live = block_sum(live, red);
        const float md = block_sum(sd, red) / float(B);
        const float mr = block_sum(sr, red) / float(B);
        float vd = 0.f, vr = 0.f;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            vd = fmaf(dv[b] - md, dv[b] - md, vd);
            vr = fmaf(rv[b] - mr, rv[b] - mr, vr);
        }
        vd = block_sum(vd, red);
        vr = block_sum(vr, red);
        if (threadIdx.x == 0) {
            const float scale = float(B) / live;  // live == 0: inf / nan rows, dropped by the caller (main.py:807-809)
            const float den = float(B > 1 ? B - 1 : 1);
            float* o = table + (long long)t * 5;
            o[0] = md * scale; o[1] = sqrtf(vd / den) * scale; o[2] = mr * scale; o[3] = sqrtf(vr / den) * scale;
            o[4] = live;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------
// Masked mean-squared error of the reward predictions (reference main.py:182-186), forward and gradient in one
// single-block launch:  loss = scale * (*scale_dev) / (B*R) * sum_b mask[b] * sum_r (pred - target)^2
// scale_dev (optional device scalar) carries the training-progress factor theta = iter/iters of main.py:143,185 so
// that one captured CUDA graph serves every iteration; loss_raw (optional) receives the unscaled masked mean, the
// value the reference logs as "Rd Loss" (main.py:184).
// ----------------------------------------------------------------------------------------------
// Sequence form like bce_logits_kernel: pred is [T][B][R] (t-major), target / mask are [B, T] windows addressed by
// (b, t) strides; loss = scale * theta * sum_t mean_t, loss_raw[t] = the unscaled masked mean of step t.  One block
// walks the T steps in order (deterministic sum; T*B*R is a few hundred elements).
__global__ void masked_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                  long long t_bstride, long long t_tstride, const float* __restrict__ mask,
                                  long long m_bstride, long long m_tstride, int T, int B, int R, float scale,
                                  const float* __restrict__ scale_dev, float* __restrict__ loss,
                                  float* __restrict__ loss_raw, float* __restrict__ dpred) {
    pdl_sync();
    __shared__ float red[33];
    const float k0 = 1.f / (float(B) * float(R));
    const float k = scale * (scale_dev ? __ldg(scale_dev) : 1.f) * k0;
    float total = 0.f;
    for (int t = 0; t < T; ++t) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < B * R; i += blockDim.x) {
            const int b = i / R, r = i - b * R;
            const float m = mask ? __ldg(mask + (long long)b * m_bstride + (long long)t * m_tstride) : 1.f;
            const float d = pred[(long long)t * B * R + i] -
                            __ldg(target + (long long)b * t_bstride + (long long)t * t_tstride + r);
            acc = fmaf(m * d, d, acc);
            if (dpred) dpred[(long long)t * B * R + i] = 2.f * k * m * d;
        }
        acc = block_sum(acc, red);
        total += acc;
        if (threadIdx.x == 0 && loss_raw) loss_raw[t] = acc * k0;
    }
    if (threadIdx.x == 0) *loss = total * k;
}

// ----------------------------------------------------------------------------------------------
// Fused clip_grad_value_ + Adam (reference main.py:287-296; torch.optim.Adam defaults, no weight decay).
// Multi-tensor: one launch updates every parameter of every network.
// ----------------------------------------------------------------------------------------------
struct AdamChunk {
    float* p;
    const float* g;
    float* m;
    float* v;
    int n;
    float clip;  // <= 0: no clipping
    const float* step;  // per-chunk device step counter (torch keeps one per parameter), or nullptr
};
constexpr int kAdamChunksPerLaunch = 48;
struct AdamArgs {
    AdamChunk chunk[kAdamChunksPerLaunch];
    int count;
    float lr, beta1, beta2, eps;
    float bc1, bc2_sqrt;  // 1 - beta1^t, sqrt(1 - beta2^t) (ignored when step_ptr != nullptr)
    const float* step_ptr;  // device step counter (float) for graph replay, or nullptr
    float gscale;           // multiplies the gradient before clipping (1/world_size for DP)
};

__global__ void clip_adam_kernel(const __grid_constant__ AdamArgs A) {
    pdl_sync();
    const AdamChunk& C = A.chunk[blockIdx.y];
    float bc1 = A.bc1, bc2s = A.bc2_sqrt;
    const float* sp = C.step ? C.step : A.step_ptr;
    if (sp) {
        const float t = __ldg(sp);
        bc1 = 1.f - powf(A.beta1, t);
        bc2s = sqrtf(1.f - powf(A.beta2, t));
    }
    const float step_size = A.lr / bc1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.n; i += gridDim.x * blockDim.x) {
        float g = C.g[i] * A.gscale;
        if (C.clip > 0.f) g = fminf(fmaxf(g, -C.clip), C.clip);
        const float m = A.beta1 * C.m[i] + (1.f - A.beta1) * g;
        const float v = A.beta2 * C.v[i] + (1.f - A.beta2) * g * g;
        C.m[i] = m;
        C.v[i] = v;
        const float denom = sqrtf(v) / bc2s + A.eps;
        C.p[i] -= step_size * (m / denom);
    }
}

}  // namespace scm

namespace scm {

// ----------------------------------------------------------------------------------------------
// Reward head (reference models.py:240-250): the second conv is stride-2 / valid, evaluated here as a
// stride-1 same-size conv whose outputs are only consumed on the lattice (2a+2, 2b+2) of interior
// coordinates.  y2: fp32 [B][3R][H][W] (class-major channels: k*R + j).
//   r[b][j] = sum_{lattice} softmax_k(y2)[0] - softmax_k(y2)[2]
// Optionally writes the per-pixel map [B][R][h2][w2] (visualize=True).
// ----------------------------------------------------------------------------------------------
__global__ void reward_head_fwd_kernel(const float* __restrict__ y2, int B, int R, int H, int W, int h2, int w2,
                                       float* __restrict__ r, float* __restrict__ map) {
    pdl_sync();
    __shared__ float red[33];
    const int b = blockIdx.x, j = blockIdx.y;
    const size_t hw = size_t(H) * W;
    const float* base = y2 + (size_t(b) * 3 * R + j) * hw;
    float acc = 0.f;
    for (int i = threadIdx.x; i < h2 * w2; i += blockDim.x) {
        const int a = i / w2, c = i - a * w2;
        const size_t pos = size_t(2 * a + 2) * W + (2 * c + 2);
        const float l0 = __ldg(base + pos), l1 = __ldg(base + size_t(R) * hw + pos),
                    l2 = __ldg(base + size_t(2 * R) * hw + pos);
        const float m = fmaxf(l0, fmaxf(l1, l2));
        const float e0 = __expf(l0 - m), e1 = __expf(l1 - m), e2 = __expf(l2 - m);
        const float d = (e0 - e2) / (e0 + e1 + e2);
        if (map) map[(size_t(b) * R + j) * h2 * w2 + i] = d;
        acc += d;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) r[b * R + j] = acc;
}

// d2 plane [B][H+2][W+2][16] (bf16): zero everywhere except lattice pixels, where for class k of reward j
//   d logit_k = dr[b][j] * p_k * ((k==0) - (k==2) - (p_0 - p_2))
__global__ void reward_head_bwd_kernel(const float* __restrict__ y2, const float* __restrict__ dr, int B, int R, int H,
                                       int W, int h2, int w2, __nv_bfloat16* __restrict__ d2) {
    pdl_sync();
    const int Hp = H + 2, Wp = W + 2;
    const long long rows = (long long)B * Hp * Wp;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows) return;
    const int b = int(p / (Hp * Wp));
    const int rem = int(p - (long long)b * Hp * Wp);
    const int hp = rem / Wp, wp = rem - hp * Wp;
    const int h = hp - 1, w = wp - 1;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    const bool on = h >= 2 && w >= 2 && ((h & 1) == 0) && ((w & 1) == 0) && (h - 2) / 2 < h2 && (w - 2) / 2 < w2;
    if (on) {
        const size_t hw = size_t(H) * W;
        const size_t pos = size_t(h) * W + w;
        for (int j = 0; j < R; ++j) {
            const float* base = y2 + (size_t(b) * 3 * R + j) * hw + pos;
            const float l0 = __ldg(base), l1 = __ldg(base + size_t(R) * hw), l2 = __ldg(base + size_t(2 * R) * hw);
            const float m = fmaxf(l0, fmaxf(l1, l2));
            const float e0 = __expf(l0 - m), e1 = __expf(l1 - m), e2 = __expf(l2 - m);
            const float inv = 1.f / (e0 + e1 + e2);
            const float p0 = e0 * inv, p1 = e1 * inv, p2 = e2 * inv;
            const float d = p0 - p2, g = __ldg(dr + b * R + j);
            v[j] = g * p0 * (1.f - d);
            v[R + j] = g * p1 * (-d);
            v[2 * R + j] = g * p2 * (-1.f - d);
        }
    }
    uint4 o0, o1;
    __nv_bfloat162* w0 = reinterpret_cast<__nv_bfloat162*>(&o0);
    __nv_bfloat162* w1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w0[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w1[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
    }
    uint4* op = reinterpret_cast<uint4*>(d2 + p * 16);
    op[0] = o0;
    op[1] = o1;
}

}  // namespace scm

namespace scm {

// ----------------------------------------------------------------------------------------------
// Counterfactual losses (reference main.py:258-262 and 279-283), fused abs-diff / mean / mask reductions.
//   za, zb: fp32 [B][L][H][W].  rowmean[b][l] = mean_{h,w} |za - zb|   (written for the backward pass)
//   mode 0 (disentanglement): loss += lambda/B * mask[b] * (1/L) * sum_l rowmean[b][l] * unswapped[b][l]
//   mode 1 (action control):  loss += lambda/B * mask[b] * -log( (1/L) * sum_l rowmean[b][l] + 1e-3 )
// Two launches: one block per (b, l) row mean (B*L blocks keep the whole chip busy; a single block per sample walking
// its L rows took 129 us at 32 x 16 x 64 x 64), then one block combines the rows of every sample.
// ----------------------------------------------------------------------------------------------
__global__ void cf_rowmean_kernel(const float* __restrict__ za, const float* __restrict__ zb, int HW,
                                  float* __restrict__ rowmean) {
    pdl_sync();
    __shared__ float red[33];
    const size_t row = blockIdx.x;  // b * L + l
    const float4* pa = reinterpret_cast<const float4*>(za + row * HW);
    const float4* pb = reinterpret_cast<const float4*>(zb + row * HW);
    float acc = 0.f;
    if ((HW & 3) == 0) {
        for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
            const float4 a = __ldg(pa + i), c = __ldg(pb + i);
            acc += fabsf(a.x - c.x) + fabsf(a.y - c.y) + fabsf(a.z - c.z) + fabsf(a.w - c.w);
        }
    } else {
        const float* qa = za + row * HW;
        const float* qb = zb + row * HW;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) acc += fabsf(__ldg(qa + i) - __ldg(qb + i));
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) rowmean[row] = acc / float(HW);
}

__global__ void cf_loss_fwd_kernel(const float* __restrict__ rowmean, const float* __restrict__ unswapped,
                                   const float* __restrict__ mask, int B, int L, int mode, float lambda,
                                   float* __restrict__ loss) {
    pdl_sync();
    __shared__ float red[33];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += rowmean[b * L + l] * (mode == 0 ? __ldg(unswapped + b * L + l) : 1.f);
        s /= float(L);
        const float term = (mode == 0) ? s : -logf(s + 1e-3f);
        acc += lambda / float(B) * __ldg(mask + b) * term;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) *loss += acc;
}

// d loss / d za (and the negative for zb): gscale * coef[b][l] * sign(za - zb) / HW
//   mode 0: coef = lambda/B * mask[b] * unswapped[b][l] / L
//   mode 1: coef = -lambda/B * mask[b] / (L * (mean_l rowmean[b][l] + 1e-3))
__global__ void cf_loss_bwd_kernel(const float* __restrict__ za, const float* __restrict__ zb,
                                   const float* __restrict__ unswapped, const float* __restrict__ mask,
                                   const float* __restrict__ rowmean, const float* __restrict__ gscale, int B, int L,
                                   int HW, int mode, float lambda, float* __restrict__ dza, float* __restrict__ dzb) {
    pdl_sync();
    const int b = blockIdx.y, l = blockIdx.z;
    float coef;
    if (mode == 0) {
        coef = lambda / float(B) * __ldg(mask + b) * __ldg(unswapped + b * L + l) / float(L);
    } else {
        float s = 0.f;
        for (int j = 0; j < L; ++j) s += __ldg(rowmean + b * L + j);
        s /= float(L);
        coef = -lambda / float(B) * __ldg(mask + b) / (float(L) * (s + 1e-3f));
    }
    coef *= __ldg(gscale) / float(HW);
    const size_t base = (size_t(b) * L + l) * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const float d = __ldg(za + base + i) - __ldg(zb + base + i);
        const float g = d > 0.f ? coef : (d < 0.f ? -coef : 0.f);
        if (dza) dza[base + i] = g;
        if (dzb) dzb[base + i] = -g;
    }
}

// ----------------------------------------------------------------------------------------------
// Stand-alone Transition tail (reference models.py:103-112): p = sigmoid(x); z = (u < p) in training or (p > 0.5)
// in eval.  The hot path fuses this into the conv6 epilogue; exported for callers that hold fp32 logits.
// ----------------------------------------------------------------------------------------------
__global__ void transition_tail_kernel(const float* __restrict__ x, const float* __restrict__ u, long long n,
                                       float* __restrict__ p, float* __restrict__ z) {
    pdl_sync();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float pv = 1.f / (1.f + __expf(-__ldg(x + i)));
        if (p) p[i] = pv;
        z[i] = u ? (__ldg(u + i) < pv ? 1.f : 0.f) : (pv > 0.5f ? 1.f : 0.f);
    }
}

// uniforms of the Philox stream as a tensor: thread t evaluates block (offset + t) once and writes its four outputs
__global__ void philox_fill_kernel(float* __restrict__ out, long long n, const unsigned long long* __restrict__ rng) {
    pdl_sync();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = t * 4;
    if (i0 >= n) return;
    const unsigned long long seed = __ldg(rng), ctr = __ldg(rng + 1) + (unsigned long long)t;
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
    const float s = 1.0f / 16777216.0f;
    const float4 u = make_float4((float(c0 >> 8) + 0.5f) * s, (float(c1 >> 8) + 0.5f) * s,
                                 (float(c2 >> 8) + 0.5f) * s, (float(c3 >> 8) + 0.5f) * s);
    if (i0 + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        reinterpret_cast<float4*>(out)[t] = u;
    } else {
        const float v[4] = {u.x, u.y, u.z, u.w};
        for (int j = 0; j < 4 && i0 + j < n; ++j) out[i0 + j] = v[j];
    }
}

// ----------------------------------------------------------------------------------------------
// Device-side replay buffer: the sampler of the reference's get_trajectories (envs/minipacman.py:122-164) as one
// kernel.  Episodes live in HBM in fixed slots ([slots][max_len] frames / rewards / actions, ep_len[slot] valid steps;
// the first *n_filled slots hold an episode).  Row b of the batch is a concatenation of clips:
//     while remaining > 0:  ep = random.choice(buffer); start = randint(0, len - 3) (or 0); end = min(start + remaining,
//                           len - 1); take ep[start:end]; dones = [False]*(duration-1) + [True]; remaining -= duration
// Clip k of row b draws its two uniforms from Philox counter (offset + b*Hn + k), lanes 0 / 1 - i.e. elements
// 4*(b*Hn+k) and 4*(b*Hn+k)+1 of the scmgan_philox_uniform stream of the same state - so the plan can be re-derived
// outside (tests) and does not depend on the launch geometry.  One block per batch row: thread 0 lays out the row
// (which slot / source step every timestep comes from), then the block copies the frames with 128-bit accesses.
// ----------------------------------------------------------------------------------------------
struct ReplayParams {
    const float* frames;   // [slots][max_len][per_frame]
    const float* rewards;  // [slots][max_len][R]
    const int* actions;    // [slots][max_len]
    const int* ep_len;     // [slots]
    const int* n_filled;   // device scalar
    int slots, max_len, R;
    long long per_frame;
    int B, Hn, random_start;
    const unsigned long long* rng;  // {seed, offset}
    float* states;         // [B][Hn][per_frame]
    float* rewards_out;    // [B][Hn][R]
    float* dones;          // [B][Hn]
    long long* actions_out;  // [B][Hn]
    int* plan;             // optional [B][Hn][3]: (slot, start, duration) of clip k, -1 past the last clip
};

__global__ void replay_sample_kernel(const ReplayParams P) {
    pdl_sync();
    extern __shared__ int rp_smem[];  // [Hn] source slot, [Hn] source step, [Hn] done flag
    int* src_slot = rp_smem;
    int* src_step = src_slot + P.Hn;
    int* src_done = src_step + P.Hn;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        const unsigned long long seed = __ldg(P.rng), off = __ldg(P.rng + 1);
        const int nf = max(1, min(__ldg(P.n_filled), P.slots));
        int remaining = P.Hn, pos = 0, k = 0;
        while (remaining > 0) {
            const unsigned long long ctr = off + (unsigned long long)b * P.Hn + k;
            const float u0 = philox_uniform(seed, ctr, 0), u1 = philox_uniform(seed, ctr, 1);
            const int slot = min(int(u0 * float(nf)), nf - 1);
            const int len = __ldg(P.ep_len + slot);
            int start = 0;
            if (P.random_start) start = min(int(u1 * float(len - 3)), len - 4);  // np.random.randint(0, len - 3)
            start = max(start, 0);
            const int end = min(start + remaining, len - 1);
            const int dur = max(end - start, 1);  // an episode shorter than 2 steps would never terminate the loop
            for (int i = 0; i < dur && pos < P.Hn; ++i, ++pos) {
                src_slot[pos] = slot;
                src_step[pos] = min(start + i, len - 1);
                src_done[pos] = (i == dur - 1);
            }
            if (P.plan) {
                int* q = P.plan + ((long long)b * P.Hn + k) * 3;
                q[0] = slot; q[1] = start; q[2] = dur;
            }
            remaining -= dur;
            ++k;
        }
        if (P.plan)
            for (; k < P.Hn; ++k) {
                int* q = P.plan + ((long long)b * P.Hn + k) * 3;
                q[0] = -1; q[1] = -1; q[2] = -1;
            }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < P.Hn; t += blockDim.x) {
        const long long src = (long long)src_slot[t] * P.max_len + src_step[t];
        const long long dst = (long long)b * P.Hn + t;
        P.dones[dst] = src_done[t] ? 1.f : 0.f;
        P.actions_out[dst] = (long long)__ldg(P.actions + src);
        for (int r = 0; r < P.R; ++r) P.rewards_out[dst * P.R + r] = __ldg(P.rewards + src * P.R + r);
    }
    const bool vec = (P.per_frame % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.frames) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(P.states) & 15) == 0);
    for (int t = 0; t < P.Hn; ++t) {
        const float* s = P.frames + ((long long)src_slot[t] * P.max_len + src_step[t]) * P.per_frame;
        float* d = P.states + ((long long)b * P.Hn + t) * P.per_frame;
        if (vec) {
            const float4* s4 = reinterpret_cast<const float4*>(s);
            float4* d4 = reinterpret_cast<float4*>(d);
            for (long long i = threadIdx.x; i < P.per_frame / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
        } else {
            for (long long i = threadIdx.x; i < P.per_frame; i += blockDim.x) d[i] = __ldg(s + i);
        }
    }
}

// advance the device-side Philox offset after a sampling launch (keeps the whole step CUDA-graph replayable)
__global__ void rng_advance_kernel(unsigned long long* rng, unsigned long long n) {
    pdl_sync(); rng[1] += n; }

}  // namespace scm
