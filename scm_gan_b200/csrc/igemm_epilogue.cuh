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

// Compiler-only fence: ties 16 accumulator registers to a point after tcgen05.wait::ld so that no arithmetic on them
// can be scheduled above the wait (the asynchronous tcgen05.ld has only then delivered them).
__device__ __forceinline__ void epi_reg_fence16(float* v) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                      "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]),
                      "+f"(v[15]));
}

__device__ __forceinline__ void epi_reg_fence8(float* v) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
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

// GROUP: columns per pass (32; 16 when the CTA's N slice is not a multiple of 32; 8 for the 16-column fp32 heads so
// that both warps of a lane quarter have work - only without a bf16 plane output, whose stores are 16 columns wide).
// F32: the fp32 NCHW / Bernoulli-head output is compiled in (only the narrow N = 16 heads use it; keeping it out of
// the wide instantiations keeps the hot loop small enough for the instruction cache).
template <int GROUP, bool F32>
__device__ __forceinline__ void igemm_epilogue_tile(const IgemmParams& P, int n_total, int n0, int ncols_cta, int p,
                                                    int half, int lane, uint32_t taddr, const float* s_bias) {
    static_assert(GROUP == 8 || GROUP == 16 || GROUP == 32, "unsupported column group");
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
    const uint32_t s_bias_u32 = smem_u32(s_bias);
    __nv_bfloat16* const ob = P.out + (P.out_c_off + n0);
    __nv_bfloat16* const ob0 = ob + size_t(d0 < 0 ? 0 : d0) * P.out_cs;
    const bool any_wrap = (d1 >= 0) || (d2 >= 0);

    // TMEM reads are double-buffered: the tcgen05.ld of the warp's next column group is in flight while the current
    // group goes through the arithmetic and the stores (the epilogue is latency-bound, not issue-bound).
    auto load = [&](float (&v)[GROUP], int c0) {
        if constexpr (GROUP == 8) {
            tmem_ld8(taddr + uint32_t(c0), v);
        } else {
            tmem_ld16(taddr + uint32_t(c0), v);
            if constexpr (GROUP == 32) tmem_ld16(taddr + uint32_t(c0 + 16), v + 16);
        }
    };
    auto process = [&](float (&v)[GROUP], int c0) {
        if constexpr (GROUP == 8) {
            epi_reg_fence8(v);
        } else {
            epi_reg_fence16(v);
            if constexpr (GROUP == 32) epi_reg_fence16(v + 16);
        }
        if (P.debug & 1) return;
        {
#pragma unroll
            for (int j = 0; j < GROUP / 4; ++j) {
                float4 b4;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                             : "r"(s_bias_u32 + uint32_t(c0 + 4 * j) * 4u));
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
                const int co = c0 + j * 16;
                if (!(P.debug & 64)) {
                    if (d0 >= 0) st_global_256(ob0 + co, o);
                    if (any_wrap) {  // border rows only (divergent, rare)
                        if (d1 >= 0) st_global_256(ob + size_t(d1) * P.out_cs + co, o);
                        if (d2 >= 0) st_global_256(ob + size_t(d2) * P.out_cs + co, o);
                        if (d3 >= 0) st_global_256(ob + size_t(d3) * P.out_cs + co, o);
                    }
                }
            }
        }
        if (F32 && P.out_f32 && R.interior) {
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
    };
    float va[GROUP], vb[GROUP];
    int c0 = half * GROUP;
    if (c0 < ncols_cta) load(va, c0);
    while (c0 < ncols_cta) {
        tmem_ld_wait();
        const int c1 = c0 + 2 * GROUP;
        if (c1 < ncols_cta) load(vb, c1);
        process(va, c0);
        if (c1 >= ncols_cta) break;
        tmem_ld_wait();
        c0 = c1 + 2 * GROUP;
        if (c0 < ncols_cta) load(va, c0);
        process(vb, c1);
    }
}

}  // namespace scm
