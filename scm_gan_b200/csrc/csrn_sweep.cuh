// CSRN directional sweep (reference spatial_recurrent.py:61-114): one pass over the image lines in a fixed order.
// Per line i (n pixels, C channels):
//     out_i  = GRU_step(x_i, s)          single-step nn.GRU(C, C, bias=False), gates r, z, n (torch order)
//     ctx_i  = out_i                     the line's context
//     s      = tanh(Conv1d_k3,p1(out_i) + b)   hidden state handed to the next line
// The sweep is strictly sequential over lines and independent across batch samples: one persistent CTA per sample
// keeps the line buffers in shared memory and walks the lines; everything is fp32 CUDA-core arithmetic (an
// interface-only layer of the reference with 2(H+W) dependent steps of a few hundred KFLOP - nothing for the tensor
// core).  The reference launches a cuDNN RNN step, a conv1d, a tanh and four permute/copy kernels per line.
//
// The backward kernel walks the lines in reverse, re-computes the gates from the saved hidden states, and writes the
// weight gradients of its sample into a per-sample partial buffer (summed over samples by the caller: deterministic).
#pragma once
#include <cuda_runtime.h>

namespace scm {

struct CsrnSweepParams {
    const float* x;        // [B][C][H][W] (any layout described by the strides below)
    long long xs_b, xs_c, xs_line, xs_pix;
    int B, C, L, n;        // lines, pixels per line
    int reverse;           // walk lines L-1 .. 0
    const float* w_ih;     // [3C][C]
    const float* w_hh;     // [3C][C]
    const float* conv_w;   // [C][C][3]
    const float* conv_b;   // [C]
    float* ctx;            // forward: out, same strides as x.  backward: d ctx (input)
    float* states;         // [B][L][n][C]: hidden state ENTERING line i (forward writes, backward reads)
    // backward only
    const float* ctx_fwd;  // forward contexts (= out_i), same strides as x
    float* dx;             // same strides as x
    float* dparams;        // [B][6*C*C + 3*C*C + C] per-sample partials: dW_ih, dW_hh, dconv_w, dconv_b (zeroed)
};

__device__ __forceinline__ float csrn_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(256) csrn_sweep_fwd_kernel(const CsrnSweepParams P) {
    pdl_sync();
    extern __shared__ float sm[];
    const int C = P.C, n = P.n, nC = n * C;
    float* xs = sm;            // [n][C]
    float* st = xs + nC;       // [n][C] hidden state
    float* outs = st + nC;     // [n][C]
    const int b = blockIdx.x;
    const float* xb = P.x + (long long)b * P.xs_b;
    float* cb = P.ctx + (long long)b * P.xs_b;
    for (int i = threadIdx.x; i < nC; i += blockDim.x) st[i] = 0.f;
    __syncthreads();
    for (int step = 0; step < P.L; ++step) {
        const int line = P.reverse ? P.L - 1 - step : step;
        for (int i = threadIdx.x; i < nC; i += blockDim.x) {
            const int j = i / C, c = i - j * C;
            xs[i] = xb[c * P.xs_c + (long long)line * P.xs_line + (long long)j * P.xs_pix];
            if (P.states) P.states[(((long long)b * P.L + line) * n + j) * C + c] = st[i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nC; i += blockDim.x) {
            const int j = i / C, c = i - j * C;
            const float* xr = xs + j * C;
            const float* sr = st + j * C;
            float gi[3] = {0.f, 0.f, 0.f}, gh[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < C; ++k) {
                const float xv = xr[k], sv = sr[k];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    gi[g] = fmaf(xv, __ldg(P.w_ih + (long long)(g * C + c) * C + k), gi[g]);
                    gh[g] = fmaf(sv, __ldg(P.w_hh + (long long)(g * C + c) * C + k), gh[g]);
                }
            }
            const float r = csrn_sigmoid(gi[0] + gh[0]);
            const float z = csrn_sigmoid(gi[1] + gh[1]);
            const float nn = tanhf(gi[2] + r * gh[2]);
            const float o = (1.f - z) * nn + z * sr[c];
            outs[i] = o;
            cb[c * P.xs_c + (long long)line * P.xs_line + (long long)j * P.xs_pix] = o;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nC; i += blockDim.x) {
            const int j = i / C, co = i - j * C;
            float acc = __ldg(P.conv_b + co);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int jj = j + k - 1;
                if (jj < 0 || jj >= n) continue;
                const float* orow = outs + jj * C;
                for (int ci = 0; ci < C; ++ci) acc = fmaf(orow[ci], __ldg(P.conv_w + ((long long)co * C + ci) * 3 + k), acc);
            }
            st[i] = tanhf(acc);  // the GRU of this step has finished reading st (barrier above)
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) csrn_sweep_bwd_kernel(const CsrnSweepParams P) {
    pdl_sync();
    extern __shared__ float sm[];
    const int C = P.C, n = P.n, nC = n * C;
    float* xs = sm;             // x_i
    float* st = xs + nC;        // s_i (state entering the line)
    float* outs = st + nC;      // out_i
    float* dout = outs + nC;    // d out_i
    float* ds = dout + nC;      // d (state produced by this line), carried from the later line
    float* dpre = ds + nC;      // d (conv pre-activation)
    float* dgi = dpre + nC;     // [n][3C]
    float* dgh = dgi + 3 * nC;  // [n][3C]
    const int b = blockIdx.x;
    const float* xb = P.x + (long long)b * P.xs_b;
    const float* ob = P.ctx_fwd + (long long)b * P.xs_b;
    const float* dcb = P.ctx + (long long)b * P.xs_b;
    float* dxb = P.dx + (long long)b * P.xs_b;
    float* dw_ih = P.dparams + (long long)b * (9LL * C * C + C);
    float* dw_hh = dw_ih + 3LL * C * C;
    float* dcw = dw_hh + 3LL * C * C;
    float* dcbias = dcw + 3LL * C * C;
    for (int i = threadIdx.x; i < nC; i += blockDim.x) ds[i] = 0.f;
    __syncthreads();
    for (int step = P.L - 1; step >= 0; --step) {
        const int line = P.reverse ? P.L - 1 - step : step;
        const int next_line = P.reverse ? line - 1 : line + 1;  // the line that consumed this line's state
        for (int i = threadIdx.x; i < nC; i += blockDim.x) {
            const int j = i / C, c = i - j * C;
            const long long off = c * P.xs_c + (long long)line * P.xs_line + (long long)j * P.xs_pix;
            xs[i] = xb[off];
            outs[i] = ob[off];
            dout[i] = dcb[off];
            st[i] = P.states[(((long long)b * P.L + line) * n + j) * C + c];
            float dp = 0.f;
            if (step < P.L - 1) {
                const float sn = P.states[(((long long)b * P.L + next_line) * n + j) * C + c];
                dp = ds[i] * (1.f - sn * sn);
            }
            dpre[i] = dp;
        }
        __syncthreads();
        if (step < P.L - 1) {
            // through s_next = tanh(conv1d(out) + b): d out, d conv weight, d conv bias
            for (int i = threadIdx.x; i < nC; i += blockDim.x) {
                const int j = i / C, ci = i - j * C;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int jp = j - k + 1;  // pre[jp] read out[jp + k - 1] = out[j]
                    if (jp < 0 || jp >= n) continue;
                    const float* drow = dpre + jp * C;
                    for (int co = 0; co < C; ++co) acc = fmaf(drow[co], __ldg(P.conv_w + ((long long)co * C + ci) * 3 + k), acc);
                }
                dout[i] += acc;
            }
            for (int i = threadIdx.x; i < 3 * C * C; i += blockDim.x) {
                const int k = i % 3, ci = (i / 3) % C, co = i / (3 * C);
                float acc = 0.f;
                for (int j = 0; j < n; ++j) {
                    const int jj = j + k - 1;
                    if (jj < 0 || jj >= n) continue;
                    acc = fmaf(dpre[j * C + co], outs[jj * C + ci], acc);
                }
                dcw[i] += acc;
            }
            for (int co = threadIdx.x; co < C; co += blockDim.x) {
                float acc = 0.f;
                for (int j = 0; j < n; ++j) acc += dpre[j * C + co];
                dcbias[co] += acc;
            }
        }
        __syncthreads();
        // GRU backward at this line (gates re-computed)
        for (int i = threadIdx.x; i < nC; i += blockDim.x) {
            const int j = i / C, c = i - j * C;
            const float* xr = xs + j * C;
            const float* sr = st + j * C;
            float gi[3] = {0.f, 0.f, 0.f}, gh[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < C; ++k) {
                const float xv = xr[k], sv = sr[k];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    gi[g] = fmaf(xv, __ldg(P.w_ih + (long long)(g * C + c) * C + k), gi[g]);
                    gh[g] = fmaf(sv, __ldg(P.w_hh + (long long)(g * C + c) * C + k), gh[g]);
                }
            }
            const float r = csrn_sigmoid(gi[0] + gh[0]);
            const float z = csrn_sigmoid(gi[1] + gh[1]);
            const float nn = tanhf(gi[2] + r * gh[2]);
            const float d_o = dout[i];
            const float d_nn = d_o * (1.f - z);
            const float d_z = d_o * (sr[c] - nn);
            const float d_an = d_nn * (1.f - nn * nn);
            const float d_ar = d_an * gh[2] * r * (1.f - r);
            const float d_az = d_z * z * (1.f - z);
            float* gi_o = dgi + j * 3 * C;
            float* gh_o = dgh + j * 3 * C;
            gi_o[c] = d_ar; gi_o[C + c] = d_az; gi_o[2 * C + c] = d_an;
            gh_o[c] = d_ar; gh_o[C + c] = d_az; gh_o[2 * C + c] = d_an * r;
            dpre[i] = d_o * z;  // direct path out -> s (re-uses dpre as the new ds accumulator)
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nC; i += blockDim.x) {
            const int j = i / C, ci = i - j * C;
            const float* gi_r = dgi + j * 3 * C;
            const float* gh_r = dgh + j * 3 * C;
            float ax = 0.f, as = 0.f;
            for (int gc = 0; gc < 3 * C; ++gc) {
                ax = fmaf(gi_r[gc], __ldg(P.w_ih + (long long)gc * C + ci), ax);
                as = fmaf(gh_r[gc], __ldg(P.w_hh + (long long)gc * C + ci), as);
            }
            dxb[ci * P.xs_c + (long long)line * P.xs_line + (long long)j * P.xs_pix] = ax;
            ds[i] = dpre[i] + as;
        }
        for (int i = threadIdx.x; i < 3 * C * C; i += blockDim.x) {
            const int ci = i % C, gc = i / C;
            float a1 = 0.f, a2 = 0.f;
            for (int j = 0; j < n; ++j) {
                a1 = fmaf(dgi[j * 3 * C + gc], xs[j * C + ci], a1);
                a2 = fmaf(dgh[j * 3 * C + gc], st[j * C + ci], a2);
            }
            dw_ih[i] += a1;
            dw_hh[i] += a2;
        }
        __syncthreads();
    }
}

}  // namespace scm
