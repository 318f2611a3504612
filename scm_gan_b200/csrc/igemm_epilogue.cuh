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

// Fused decoder loss head: per-warp running sum of the current rollout step's loss share.  Every epilogue warp owns one
// row [T] of a scratch table; a step's share is added to it when the warp moves on to another step (tiles of one warp
// visit the steps in increasing order) and at the end of the kernel, and bce_finalize_kernel sums the rows in a fixed
// order: no atomics, the same bits on every run.
struct BceCarry {
    float* row;  // this warp's scratch row [T]
    int t;       // step of the running sum (-1: none)
    float sum;
    __device__ __forceinline__ void add(int step, float v, int lane) {
        if (step != t) {
            flush(lane);
            t = step;
        }
        sum += v;
    }
    __device__ __forceinline__ void flush(int lane) {
        if (t >= 0 && lane == 0) row[t] += sum;
        sum = 0.f;
    }
};

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
// The warp's tile is final when `acc_bar` completes phase `acc_phase`; everything that does not need the accumulator
// (row decode, and the LeakyReLU-derivative gate of dgrad: the sign bits of this row's slice of the gate plane, 128
// columns per warp at most -> four registers) is fetched BEFORE that wait, so that the global-memory latency of the
// gate plane (a saved forward activation, usually in HBM) overlaps the MMAs instead of stalling every column group.
// SHARE: epilogue warps per TMEM lane quarter (2, or 4 in the narrow-K kernel whose short main loop leaves the
// epilogue exposed); the warps of a quarter interleave the column groups.
// BCE: fused decoder loss head (IgemmParams::bce_*; only with GROUP = 16, a plane output and no activation): the group's
// 16 columns are the (padded) colour channels of one pixel, v becomes d loss / d logit and the warp adds its share of the
// masked BCE mean of the tile's rollout step to its BceCarry.
template <int GROUP, bool F32, int SHARE = 2, bool BCE = false>
__device__ __forceinline__ void igemm_epilogue_tile(const IgemmParams& P, int n_total, int n0, int ncols_cta, int p,
                                                    int half, int lane, uint32_t taddr, const float* s_bias,
                                                    uint64_t* acc_bar, uint32_t acc_phase, int p_next,
                                                    BceCarry* bce = nullptr) {
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
    if (P.gate && p_next >= 0 && p_next < P.rows) {
        // pull the NEXT tile's gate row into L2 while this tile is processed (the plane is a saved forward
        // activation that normally sits in HBM); costs no registers
        const char* gn = reinterpret_cast<const char*>(P.gate + size_t(p_next) * P.gate_cs + P.gate_c_off + n0);
        const int bytes = ncols_cta * 2;
        for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + o));
    }
    uint32_t gb[4] = {0u, 0u, 0u, 0u};  // bit (k*GROUP + i): gate of column i of this warp's k-th group is "positive"
    if (P.gate && R.interior) {
        // phase 1: every load of the tile in flight at once (the accumulator registers are dead here); 256-bit loads
        // because each lane reads its own row: a request costs one L1 wavefront per lane whatever its width
        constexpr int NG = 256 / (SHARE * GROUP);  // column groups per warp at most
        constexpr int WPG = GROUP / 2;       // 32-bit words per group
        uint32_t raw[NG * WPG];
        const __nv_bfloat16* grow = P.gate + size_t(p) * P.gate_cs + P.gate_c_off + n0;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const int c0 = half * GROUP + k * SHARE * GROUP;
            if (c0 < ncols_cta) {
                if constexpr (GROUP >= 16) {
#pragma unroll
                    for (int j = 0; j < GROUP / 16; ++j) {
                        uint32_t* w = raw + k * WPG + 8 * j;
                        asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                              "=r"(w[7])
                            : "l"(grow + c0 + 16 * j));
                    }
                } else {
                    const uint4 r = __ldg(reinterpret_cast<const uint4*>(grow + c0));
                    raw[k * WPG] = r.x; raw[k * WPG + 1] = r.y; raw[k * WPG + 2] = r.z; raw[k * WPG + 3] = r.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < WPG; ++q) raw[k * WPG + q] = 0u;
            }
        }
        // phase 2: per bf16 half, bit <- (value > 0), i.e. magnitude != 0 and sign clear
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            uint32_t m = 0u;
#pragma unroll
            for (int q = 0; q < WPG; ++q) {
                const uint32_t w = raw[k * WPG + q];
                const uint32_t pos = ((w & 0x7FFF7FFFu) + 0x7FFF7FFFu) & ~w & 0x80008000u;
                m |= ((pos >> 15) & 1u) << (2 * q);
                m |= (pos >> 31) << (2 * q + 1);
            }
            gb[(k * GROUP) >> 5] |= m << ((k * GROUP) & 31);
        }
    }
    // Bernoulli head: the uniforms of this warp's first column group are fetched before the accumulator wait as well.
    // Inside the store loop every load sat behind the previous (possibly aliasing) store and its own ~700-cycle
    // latency: measured 40 us of a 61 us conv6 launch.
    const size_t hw = size_t(P.H) * P.W;
    float upre[GROUP];
    bool have_upre = false;
    if constexpr (F32) {
        if (P.sample_out && P.uniforms && R.interior) {
            const size_t base = (size_t(R.b) * P.n_valid) * hw + size_t(R.hp - 1) * P.W + (R.wp - 1);
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
                const int n = n0 + half * GROUP + i;
                upre[i] = n < P.n_valid ? __ldg(P.uniforms + base + size_t(n) * hw) : 0.f;
            }
            have_upre = true;
        }
    }
    float bce_acc = 0.f, bce_m = 0.f, bce_inv = 0.f;
    int bce_t = 0;
    const float* bce_yp = nullptr;
    if constexpr (BCE) {
        if (R.interior) {
            bce_t = R.b / P.bce_B;
            const int bb = R.b - bce_t * P.bce_B;
            bce_m = P.bce_mask ? __ldg(P.bce_mask + (long long)bb * P.bce_mbs + (long long)bce_t * P.bce_mts) : 1.f;
            bce_inv = 1.f / (float(P.bce_B) * float(P.n_valid) * float(hw));
            bce_yp = P.bce_y + (long long)bb * P.bce_ybs + (long long)bce_t * P.bce_yts + size_t(R.hp - 1) * P.W + (R.wp - 1);
        }
    }
    // the pixel's target values are fetched before the accumulator wait as well (same reason as the uniforms above)
    float ypre[BCE ? GROUP : 1];
    if constexpr (BCE) {
#pragma unroll
        for (int i = 0; i < GROUP; ++i) {
            const int n = n0 + half * GROUP + i;
            ypre[i] = (R.interior && n < P.n_valid) ? __ldg(bce_yp + size_t(n) * hw) : 0.f;
        }
    }
    mbar_wait(acc_bar, acc_phase);
    tc_fence_after();
    const float* sbp = (P.sample_bias && R.valid) ? P.sample_bias + size_t(R.b) * n_total + n0 : nullptr;
    const float rs = (P.sample_scale && R.valid) ? P.scale * __ldg(P.sample_scale + R.b) : P.scale;
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
                v[4 * j] = fmaf(v[4 * j], rs, b4.x);
                v[4 * j + 1] = fmaf(v[4 * j + 1], rs, b4.y);
                v[4 * j + 2] = fmaf(v[4 * j + 2], rs, b4.z);
                v[4 * j + 3] = fmaf(v[4 * j + 3], rs, b4.w);
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
        if constexpr (BCE) {
            // logits -> d loss / d logits (reference main.py:189-197, 310-312: sigmoid, F.binary_cross_entropy with its log
            // clamp at -100, mean over C,H,W, mask, mean over the batch); same arithmetic as bce_logits_kernel
            if (R.interior) {
                const float k = bce_m * bce_inv;
#pragma unroll
                for (int i = 0; i < GROUP; ++i) {
                    const int n = n0 + c0 + i;
                    if (n < P.n_valid) {
                        const float yv = (c0 == half * GROUP) ? ypre[i] : __ldg(bce_yp + size_t(n) * hw);
                        const float pr = 1.f / (1.f + __expf(-v[i]));
                        const float lp = fmaxf(__logf(pr), -100.f), lq = fmaxf(__logf(1.f - pr), -100.f);
                        bce_acc -= yv * lp + (1.f - yv) * lq;
                        v[i] = (pr - yv) * k;
                    } else {
                        v[i] = 0.f;
                    }
                }
            }
        }
        if (R.interior && P.gate) {
            const int bit0 = ((c0 - half * GROUP) / (SHARE * GROUP)) * GROUP;  // this group's first bit in gb
            const int wi = bit0 >> 5;
            const uint32_t word = wi == 0 ? gb[0] : (wi == 1 ? gb[1] : (wi == 2 ? gb[2] : gb[3]));
            const uint32_t m = word >> (bit0 & 31);
#pragma unroll
            for (int i = 0; i < GROUP; ++i) v[i] *= ((m >> i) & 1u) ? 1.f : P.slope;
        }
        if (P.out) {
            // direct path: every lane writes its own row with 256-bit stores (one full 32-byte sector per 16
            // columns) plus its own wrapped copies; no shared-memory round trip, no warp synchronisation
#pragma unroll
            for (int j = 0; j < GROUP / 16; ++j) {
                uint32_t o[8];
                if (R.interior) {
                    if (P.out_fmt == FMT_F16) {  // warp-uniform
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = pack2_f16(v[16 * j + 2 * i], v[16 * j + 2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = pack2_bf16(v[16 * j + 2 * i], v[16 * j + 2 * i + 1]);
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
            const bool pre = have_upre && c0 == half * GROUP;
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
                const int n = n0 + c0 + i;
                if (n < P.n_valid) {
                    const size_t idx = base + size_t(n) * hw;
                    P.out_f32[idx] = v[i];
                    if (P.sample_out)
                        P.sample_out[idx] = pre ? (upre[i] < v[i] ? 1.f : 0.f) : bernoulli_head(P, idx, v[i]);
                }
            }
        }
    };
    if constexpr (SHARE > 2) {
        // four warps per quarter: a warp rarely owns more than one group, a second register buffer would only cost
        // occupancy headroom
        float va[GROUP];
        for (int c0 = half * GROUP; c0 < ncols_cta; c0 += SHARE * GROUP) {
            load(va, c0);
            tmem_ld_wait();
            process(va, c0);
        }
    } else {
        float va[GROUP], vb[GROUP];
        int c0 = half * GROUP;
        if (c0 < ncols_cta) load(va, c0);
        while (c0 < ncols_cta) {
            tmem_ld_wait();
            const int c1 = c0 + SHARE * GROUP;
            if (c1 < ncols_cta) load(vb, c1);
            process(va, c0);
            if (c1 >= ncols_cta) break;
            tmem_ld_wait();
            c0 = c1 + SHARE * GROUP;
            if (c0 < ncols_cta) load(va, c0);
            process(vb, c1);
        }
    }
    if constexpr (BCE) {
        if (half * GROUP < ncols_cta) {   // warp-uniform: this warp owned a column group
            // warp-reduce the tile's share step by step (one step per tile, two where a tile straddles a step boundary)
            float part = R.interior ? bce_acc * bce_m * bce_inv : 0.f;
            unsigned todo = __ballot_sync(0xFFFFFFFFu, R.interior);
            while (todo) {
                const int t0 = __shfl_sync(0xFFFFFFFFu, bce_t, __ffs(todo) - 1);
                const bool mine = R.interior && bce_t == t0;
                float sum = mine ? part : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
                bce->add(t0, sum, lane);
                todo &= ~__ballot_sync(0xFFFFFFFFu, mine);
            }
        }
    }
}

}  // namespace scm
