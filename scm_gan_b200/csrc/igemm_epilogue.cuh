// Shared epilogue of the weight-stationary convolution kernels (conv_igemm_v2.cuh, conv_igemm_v3.cuh):
// TMEM accumulator -> registers -> y = act(acc*scale + bias + sample_bias + add) * lrelu'(gate)
//   -> bf16 plane (interior rows + wrapped halo copies, or zeros on halo rows of zero-padded planes)
//   -> and/or fp32 NCHW (+ Bernoulli / threshold head).
//
// One warp owns 32 plane rows (its TMEM lane quarter); two warps share a quarter and interleave GROUP-column groups.
// Every lane writes its own row with 256-bit stores (st.global.v8.b32: one full 32-byte sector per 16 channels).
// Measured on B200 (128->128 conv, B=32, 64x64): 16-byte stores 46.4 us, shared-memory-staged 64-byte row segments
// 49.1 us (the staging round trip costs more than it saves: the epilogue is instruction-latency bound), 256-bit direct
// stores 41.2 us.  Bias comes from shared memory as LDS.128, LeakyReLU is mul+max.
#pragma once
#include "conv_igemm.cuh"

namespace scm {

// 256-bit store (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&o)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                 "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}

struct EpiRow {
    int p, b, hp, wp;
    bool valid, interior;
};

__device__ __forceinline__ EpiRow epi_decode_row(const IgemmParams& P, int p) {
    EpiRow r;
    r.p = p;
    r.valid = p < P.rows;
    r.b = 0; r.hp = 0; r.wp = 0;
    if (r.valid) {
        const int plane = P.Hp * P.Wp;
        r.b = p / plane;
        const int rem = p - r.b * plane;
        r.hp = rem / P.Wp;
        r.wp = rem - r.hp * P.Wp;
    }
    r.interior = r.valid && r.hp >= 1 && r.hp <= P.H && r.wp >= 1 && r.wp <= P.W;
    return r;
}

// GROUP: columns per pass (32, or 16 when the CTA's N slice is not a multiple of 32 / shared memory is tight).
template <int GROUP>
__device__ __forceinline__ void igemm_epilogue_tile(const IgemmParams& P, int n_total, int n0, int ncols_cta, int p,
                                                    int half, int lane, uint32_t taddr, const float* s_bias) {
    constexpr int CH = GROUP / 8;       // 16-byte chunks per row segment
    const EpiRow R = epi_decode_row(P, p);
    const int plane = P.Hp * P.Wp;
    // destinations of this row in the output plane: itself and up to three wrapped halo copies
    int d0 = -1, d1 = -1, d2 = -1, d3 = -1;
    if (R.interior || (R.valid && !P.wrap)) d0 = p;  // zero-padded planes: halo rows are written (as zeros)
    if (P.wrap && R.interior) {
        int hp2 = -1, wp2 = -1;
        if (R.hp == 1) hp2 = P.H + 1; else if (R.hp == P.H) hp2 = 0;
        if (R.wp == 1) wp2 = P.W + 1; else if (R.wp == P.W) wp2 = 0;
        if (hp2 >= 0) d1 = R.b * plane + hp2 * P.Wp + R.wp;
        if (wp2 >= 0) d2 = R.b * plane + R.hp * P.Wp + wp2;
        if (hp2 >= 0 && wp2 >= 0) d3 = R.b * plane + hp2 * P.Wp + wp2;
    }
    const float* sbp = (P.sample_bias && R.valid) ? P.sample_bias + size_t(R.b) * n_total + n0 : nullptr;
    const size_t hw = size_t(P.H) * P.W;

    for (int c0 = half * GROUP; c0 < ncols_cta; c0 += 2 * GROUP) {
        float v[GROUP];
        if (!(P.debug & 2)) {
            tmem_ld16(taddr + uint32_t(c0), v);
            if (GROUP == 32) tmem_ld16(taddr + uint32_t(c0 + 16), v + (GROUP == 32 ? 16 : 0));
            tmem_ld_wait();
        }
        if (P.debug & 1) continue;
        {
            const float4* bp = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
            for (int j = 0; j < GROUP / 4; ++j) {
                const float4 b4 = bp[j];
                v[4 * j] = fmaf(v[4 * j], P.scale, b4.x);
                v[4 * j + 1] = fmaf(v[4 * j + 1], P.scale, b4.y);
                v[4 * j + 2] = fmaf(v[4 * j + 2], P.scale, b4.z);
                v[4 * j + 3] = fmaf(v[4 * j + 3], P.scale, b4.w);
            }
        }
        if (sbp) {
#pragma unroll
            for (int j = 0; j < GROUP / 4; ++j) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sbp + c0) + j);
                v[4 * j] += s4.x; v[4 * j + 1] += s4.y; v[4 * j + 2] += s4.z; v[4 * j + 3] += s4.w;
            }
        }
        if (R.interior && P.add) {
            const uint4* ap = reinterpret_cast<const uint4*>(P.add + size_t(p) * P.add_cs + P.add_c_off + n0 + c0);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const uint4 r = __ldg(ap + j);
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&r);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 * j + i] += __bfloat162float(h[i]);
            }
        }
        if (P.act == ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < GROUP; ++i) v[i] = fmaxf(v[i], v[i] * P.slope);  // 0 < slope < 1
        } else if (P.act == ACT_SIGMOID) {
#pragma unroll
            for (int i = 0; i < GROUP; ++i) v[i] = 1.f / (1.f + __expf(-v[i]));
        }
        if (R.interior && P.gate) {
            const uint4* gp = reinterpret_cast<const uint4*>(P.gate + size_t(p) * P.gate_cs + P.gate_c_off + n0 + c0);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const uint4 r = __ldg(gp + j);
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&r);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 * j + i] *= (__bfloat162float(h[i]) > 0.f) ? 1.f : P.slope;
            }
        }
        if (P.out) {
            // direct path: every lane writes its own row with 256-bit stores (one full 32-byte sector per 16
            // columns) plus its own wrapped copies; no shared-memory round trip, no warp synchronisation
#pragma unroll
            for (int j = 0; j < GROUP / 16; ++j) {
                uint32_t o[8];
                if (R.interior) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[16 * j + 2 * i], v[16 * j + 2 * i + 1]);
                        o[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = 0u;
                }
                __nv_bfloat16* ob = P.out + (P.out_c_off + n0 + c0 + j * 16);
                if (!(P.debug & 64)) {
                    if (d0 >= 0) st_global_256(ob + size_t(d0) * P.out_cs, o);
                    if (d1 >= 0) st_global_256(ob + size_t(d1) * P.out_cs, o);
                    if (d2 >= 0) st_global_256(ob + size_t(d2) * P.out_cs, o);
                    if (d3 >= 0) st_global_256(ob + size_t(d3) * P.out_cs, o);
                }
            }
        }
        if (P.out_f32 && R.interior) {
            const size_t base = (size_t(R.b) * P.n_valid) * hw + size_t(R.hp - 1) * P.W + (R.wp - 1);
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
                const int n = n0 + c0 + i;
                if (n < P.n_valid) {
                    const size_t idx = base + size_t(n) * hw;
                    P.out_f32[idx] = v[i];
                    if (P.sample_out) P.sample_out[idx] = bernoulli_head(P, idx, v[i]);
                }
            }
        }
    }
}

}  // namespace scm
