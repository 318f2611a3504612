// 3x3 implicit-GEMM convolution, second generation: weight-stationary persistent CTAs.
//
// Measured on B200 (profiles/r01_notes.md): the first kernel re-streams weights and one activation tile
// per filter tap through the L2->SM fabric (576 KB per 128-row tile, ~40 B/cycle/SM - the fabric limit) and spends
// as long again in an epilogue of half-sector stores.  This version removes both:
//   * the packed weights of the CTA's N slice stay RESIDENT in shared memory for the whole (persistent) kernel:
//     one TMA burst at start, zero weight traffic per tile.  N is split across CTAs (blockIdx.y) until the slice
//     fits (9 * Cin * n_cta * 2 bytes <= ~147 KB);
//   * one activation "super tile" per K chunk serves SEVERAL filter taps: in the flattened plane the tap (ky,kx)
//     operand is the same rows shifted by ky*(W+2)+kx, i.e. the same shared-memory tile at a row offset.  UMMA
//     descriptors address it directly (start address + 128 B per row with base_offset = 0: the tensor core applies
//     the swizzle to the absolute shared-memory address), so A traffic drops 3x (kx reuse) or ~4.5x (ky+kx reuse);
//   * eight epilogue warps (sixteen in the narrow-K instantiation) write their rows with 256-bit stores
//     (igemm_epilogue.cuh); bias lives in shared memory.
// Warp roles and the double-buffered TMEM accumulator are as in conv_igemm.cuh.
#pragma once
#include "conv_igemm.cuh"
#include "igemm_epilogue.cuh"

namespace scm {

struct IgemmV2Geom {
    int n_total;      // packed weight rows per tap (GEMM N of the whole layer)
    int n_cta;        // N slice per CTA (multiple of 16)
    int groups;       // 9 / TPG stages per K chunk and tile
    int loads;        // TMA loads per stage
    int box_rows;     // box rows of the activation tensor map
    int ld_row[9];    // per load: row delta relative to the group's first tap row
    int ld_smem[9];   // per load: byte offset inside the stage
    uint32_t a_off16[9];  // per tap of a group: operand start inside the stage, in 16-byte units
    int a_stage_bytes;   // bytes of one pipeline stage (activation super tile [+ streamed weight tiles])
    int a_part_bytes;    // v3 streamed-weights mode: offset of the weight tiles inside a stage
    int a_soft;          // CK = 16: the producer warp stages the super tile with LDG/STS instead of TMA
    int b_streamed;      // v3: 0 = weights resident for the whole kernel, 1 = the chunk's 9 weight tiles ride in each stage
    int b_tile_bytes;    // n_cta * row bytes
    int num_stages;
    int tiles_stride;    // == gridDim.x
};

// A lone warp per scheduler issues one dependent instruction every ~4 cycles, which made the 4-warp epilogue the
// bottleneck (measured); eight epilogue warps = two per TMEM lane quarter, interleaving the 32-column groups.
constexpr int kV2EpiWarps = 8;
constexpr int kV2Threads = 64 + 32 * kV2EpiWarps;
// The narrow-K instantiation (CK = 16: nine MMAs per tile) is pure epilogue - measured 26 us for a 16->128 layer whose
// output takes 6 us to write - so it runs sixteen epilogue warps, four per lane quarter.
template <int CK> struct V2Epi { static constexpr int kWarps = (CK == 16) ? 16 : 8; };
template <int CK> constexpr int v2_threads() { return 64 + 32 * V2Epi<CK>::kWarps; }

// CK: channels per K chunk (64 -> 128B swizzle, 16 -> 32B swizzle).  TPG: filter taps served by one stage.
// HEAD selects the epilogue family compiled into the instantiation, so that no variant pays for the registers (spills in
// the epilogue were measured: conv6 39 -> 69 us) and the instruction-cache footprint of the others:
//   kHeadPlane  16-bit plane output only (hidden layers, data gradients)
//   kHeadF32    fp32 NCHW outputs: logits / dz / probabilities + Bernoulli head (optionally with a plane as well)
//   kHeadBce    fused decoder loss head (scmgan_decoder_bce_fwd)
constexpr int kHeadPlane = 0, kHeadF32 = 1, kHeadBce = 2;
template <int CK, int TPG, int HEAD = kHeadPlane>
__global__ void __launch_bounds__(v2_threads<CK>(), 1)
conv3x3_igemm_v2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ IgemmParams P, const __grid_constant__ IgemmV2Geom G) {
    using Cfg = IgemmCfg<CK>;
    constexpr int RB = Cfg::kRowBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int chunks = P.cin_chunks;
    const int b_res_bytes = 9 * chunks * G.b_tile_bytes;
    uint8_t* s_b = smem;
    uint8_t* s_a = smem + ((b_res_bytes + 1023) & ~1023);
    float* s_bias = reinterpret_cast<float*>(s_a + size_t(G.num_stages) * G.a_stage_bytes);  // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
    uint64_t* full_bar = bars;                  // [kMaxStages]
    uint64_t* empty_bar = bars + kMaxStages;    // [kMaxStages]
    uint64_t* acc_full = bars + 2 * kMaxStages;       // [2]
    uint64_t* acc_empty = bars + 2 * kMaxStages + 2;  // [2]
    uint64_t* b_full = bars + 2 * kMaxStages + 4;     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * G.n_cta;  // first output channel of this CTA's slice
    constexpr int kGroups = 9 / TPG;
    constexpr int kEpiWarps = V2Epi<CK>::kWarps;
    constexpr int kShare = kEpiWarps / 4;
    constexpr bool BCE = HEAD == kHeadBce;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < G.num_stages; ++s) {
            mbar_init(&full_bar[s], (CK == 16 && G.a_soft) ? 32 : 1);  // software staging: one arrival per producer lane
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kEpiWarps);
        }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    // Everything up to pdl_sync() is CTA-local or - with P.prewait_weights (weights packed at least two launches ago, see
    // conv_igemm_v3.cuh) - the resident weight load: it overlaps the previous kernel's tail.
    const bool early_b = P.prewait_weights != 0;
    if (!early_b) pdl_sync();
    tc_fence_before();
    __syncthreads();   // barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            // resident weights: [tap][chunk] tiles of n_cta rows
            mbar_arrive_expect_tx(b_full, uint32_t(9 * chunks * G.n_cta * RB));
            for (int tap = 0; tap < 9; ++tap)
                for (int c = 0; c < chunks; ++c)
                    tma_load_2d(s_b + size_t(tap * chunks + c) * G.b_tile_bytes, &tmap_b, b_full, c * CK,
                                tap * G.n_total + n0);
        }
        __syncwarp();
    }
    if (early_b) pdl_sync();
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < G.n_cta; i += 32 * kEpiWarps) s_bias[i] = (P.bias && n0 + i < P.bias_n) ? P.bias[n0 + i] : 0.f;
    }
    __syncthreads();   // s_bias

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform registers); one
        // elected lane issues.
        int stage = 0;
        uint32_t phase = 0;
        if (CK == 16 && G.a_soft) {
            // Narrow K (Cin = 16): a plane row is 32 bytes, and a TMA box of 32-byte rows is bound by the TMA's
            // request rate (measured: 3.7k cycles per 128-row tile for 9 x 4 KB boxes).  The producer warp instead
            // copies ONE super tile (128 + 2*Wp + 2 rows, all nine taps) per tile with 16-byte cp.async, writing the
            // 32-byte-swizzled K-major layout itself (address bit 4 ^= bit 7 on the absolute shared-memory address,
            // which is also what the row-shifted UMMA descriptors assume; rows outside the tensor are zero-filled).
            // Completion is signalled by cp.async.mbarrier.arrive, so the producer runs num_stages tiles ahead and
            // no global-memory latency sits between two tiles.
            constexpr int kIt = 17;  // 17 x 32 lanes x 16 B >= 272 rows
            const int nch = G.box_rows * 2;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += G.tiles_stride) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t sa = smem_u32(s_a + size_t(stage) * G.a_stage_bytes);
                const int row0 = tile * 128 - P.Wp - 1;
#pragma unroll
                for (int it = 0; it < kIt; ++it) {
                    const int i = it * 32 + lane;
                    if (i < nch && !(P.debug & 8)) {
                        const int row = row0 + (i >> 1);
                        const bool ok = row >= 0 && row < P.rows;
                        const __nv_bfloat16* src = P.a + (ok ? size_t(row) * P.a_cs + P.a_c_off + (i & 1) * 8 : 0);
                        uint32_t ad = sa + uint32_t(i) * 16u;
                        ad ^= (ad >> 3) & 16u;
                        if (P.coord_c >= 0 && ok && (i & 1) == (P.coord_c >> 3)) {
                            // CoordConv (reference coordconv.py:10-14): this 16-byte chunk of the pixel holds the two
                            // coordinate channels.  They exist nowhere in memory - the plane carries zeros there - and
                            // are generated here, while the im2col super tile is staged: x = -1 + 2w/W, y = -1 + 2h/H on
                            // interior pixels, 0 in the zero-padding halo (the reference pads AFTER concatenating).
                            uint4 q4 = __ldg(reinterpret_cast<const uint4*>(src));
                            const int plane = P.Hp * P.Wp;
                            const int rem = row % plane;
                            const int hp = rem / P.Wp, wp = rem - hp * P.Wp;
                            if (hp >= 1 && hp <= P.H && wp >= 1 && wp <= P.W) {
                                const uint32_t cw = pack2_fmt(-1.f + 2.f * float(wp - 1) / float(P.W),
                                                              -1.f + 2.f * float(hp - 1) / float(P.H), P.a_fmt);
                                const int wi = (P.coord_c & 7) >> 1;
                                if (wi == 0) q4.x = cw; else if (wi == 1) q4.y = cw; else if (wi == 2) q4.z = cw; else q4.w = cw;
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(q4.x), "r"(q4.y),
                                         "r"(q4.z), "r"(q4.w)
                                         : "memory");
                        } else {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ad), "l"(src),
                                         "r"(ok ? 16u : 0u)
                                         : "memory");
                        }
                    }
                }
                if (P.coord_c >= 0) {
                    // mixed cp.async / st.shared staging: wait for the copies, publish both to the async proxy, arrive
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    fence_proxy_async_smem();
                    mbar_arrive(&full_bar[stage]);
                } else {
                    // every lane: one (non-incrementing) arrival on the stage's full barrier once its copies have landed
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage]))
                                 : "memory");
                }
                if (++stage == G.num_stages) { stage = 0; phase ^= 1; }
            }
        } else
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += G.tiles_stride) {
            const uint32_t tx = uint32_t(G.loads * G.box_rows * RB);
            const int m0 = tile * 128;
            for (int g = 0; g < kGroups; ++g) {
                const int tap0 = g * TPG;
                const int row0 = m0 + (tap0 / 3 - 1) * P.Wp + (tap0 % 3 - 1);
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        uint8_t* sa = s_a + size_t(stage) * G.a_stage_bytes;
                        mbar_arrive_expect_tx(&full_bar[stage], tx);
                        for (int l = 0; l < G.loads; ++l)
                            tma_load_2d(sa + G.ld_smem[l], &tmap_a, &full_bar[stage], P.a_c_off + c * CK,
                                        row0 + G.ld_row[l]);
                    }
                    __syncwarp();
                    if (++stage == G.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        // One elected lane issues every tcgen05.mma; with N = 64 an instruction retires in ~32 cycles, so the issue
        // loop is fully unrolled, descriptor arithmetic is one 64-bit add per operand, and control flow stays
        // warp-uniform so that the descriptors live in uniform registers (no per-instruction convergence loop).
        const uint32_t idesc = make_idesc_ab(128, G.n_cta, P.a_fmt, P.b_fmt, 0, 0);
        const uint64_t bdesc0 = make_smem_desc(smem_u32(s_b), 16, Cfg::kSbo, Cfg::kLayout);
        const uint64_t adesc0 = make_smem_desc(smem_u32(s_a), 16, Cfg::kSbo, Cfg::kLayout);
        const uint32_t b_tile16 = uint32_t(G.b_tile_bytes) >> 4;
        const uint32_t a_stage16 = uint32_t(G.a_stage_bytes) >> 4;
        uint32_t a_off[TPG];
#pragma unroll
        for (int t = 0; t < TPG; ++t) a_off[t] = G.a_off16[t];
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        mbar_wait(b_full, 0);
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += G.tiles_stride) {
            mbar_wait(&acc_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + uint32_t(acc * kAccStageCols);
            uint32_t accumulate = 0;
#pragma unroll 1
            for (int g = 0; g < kGroups; ++g) {
#pragma unroll 1
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&full_bar[stage], phase);
                    if (CK == 16 && G.a_soft) fence_proxy_async_smem();  // cp.async (generic proxy) -> tensor core
                    tc_fence_after();
                    const uint64_t a_st = adesc0 + uint64_t(uint32_t(stage) * a_stage16);
                    const uint64_t b_st = bdesc0 + uint64_t(uint32_t(g * TPG * chunks + c) * b_tile16);
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < TPG; ++t) {
                            if (P.debug & 4) break;  // profiling: barrier traffic only
                            const uint64_t at = a_st + uint64_t(a_off[t]);
                            const uint64_t bt = b_st + uint64_t(uint32_t(t * chunks) * b_tile16);
#pragma unroll
                            for (int k = 0; k < Cfg::kKSteps; ++k)
                                umma_f16(tmem_d, at + uint64_t(2 * k), bt + uint64_t(2 * k), idesc,
                                         (t == 0 && k == 0) ? accumulate : 1u);
                        }
                        umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++stage == G.num_stages) { stage = 0; phase ^= 1; }
                }
            }
            if (elect_one()) umma_commit(&acc_full[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------ epilogue (igemm_epilogue.cuh) ------------------------------
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int ew = warp - 2;
        const int half = ew >> 2;  // which of the kShare warps sharing this lane quarter
        int acc = 0;
        uint32_t acc_phase = 0;
        BceCarry bce{nullptr, -1, 0.f};
        if constexpr (BCE)   // fused decoder loss head: this warp's row of the (zeroed) scratch table
            bce.row = P.bce_ws + (size_t(blockIdx.y * gridDim.x + blockIdx.x) * kEpiWarps + ew) * P.bce_T;
        int it = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += G.tiles_stride, ++it) {
            const int p = tile * 128 + q * 32 + lane;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kAccStageCols);
            if constexpr (BCE) {
                // Fused decoder loss head (host side guarantees n = 16, a plane output, no activation).  The single
                // 16-column group leaves nothing to share between the warps of a lane quarter, and the head's epilogue is
                // a long dependent chain (target loads, exp / log, stores) - longer than the tile's main loop.  The
                // kShare warps of a quarter therefore take TURNS: warp `half` owns every kShare-th tile and has kShare
                // tile periods for it; the others only acknowledge the accumulator (after it is full, so that an
                // acknowledgement can never land in the previous use of the stage).
                if (it % kShare == half)
                    igemm_epilogue_tile<16, true, kShare, true>(P, G.n_total, n0, G.n_cta, p, 0, lane, taddr, s_bias,
                                                  &acc_full[acc], acc_phase, p + kShare * G.tiles_stride * 128, &bce);
                else
                    mbar_wait(&acc_full[acc], acc_phase);
            } else if constexpr (HEAD == kHeadF32) {
                if (!P.out && G.n_cta == 16)
                    igemm_epilogue_tile<8, true, kShare>(P, G.n_total, n0, G.n_cta, p, half, lane, taddr, s_bias,
                                                  &acc_full[acc], acc_phase, p + G.tiles_stride * 128);
                else
                    igemm_epilogue_tile<16, true, kShare>(P, G.n_total, n0, G.n_cta, p, half, lane, taddr, s_bias,
                                                  &acc_full[acc], acc_phase, p + G.tiles_stride * 128);
            } else {
                if ((G.n_cta & 31) == 0 && G.n_cta >= 32 * kShare)  // else 16-column groups keep more warps busy
                    igemm_epilogue_tile<32, false, kShare>(P, G.n_total, n0, G.n_cta, p, half, lane, taddr, s_bias,
                                                  &acc_full[acc], acc_phase, p + G.tiles_stride * 128);
                else
                    igemm_epilogue_tile<16, false, kShare>(P, G.n_total, n0, G.n_cta, p, half, lane, taddr, s_bias,
                                                  &acc_full[acc], acc_phase, p + G.tiles_stride * 128);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (BCE) bce.flush(lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace scm
