// 3x3 implicit-GEMM convolution, third generation: CTA pairs (tcgen05 cta_group::2).
//
// Same weight-stationary / row-shifted-tap design as conv_igemm_v2.cuh, but two SMs of a TPC work on one 256-row
// tile: each CTA streams its own 128 activation rows, holds HALF of the packed weights (N/2 rows of every
// [tap][chunk] tile) resident in shared memory, and the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 128)
// which reads A from each CTA's own shared memory and the two B halves from both.  Per SM an MMA then reads
// 4 KB (A) + 2 KB (B half) per 64 tensor cycles = 96 B/cycle of shared-memory bandwidth instead of the 192 B/cycle
// the N-split single-CTA kernel needs, and no activation tile is loaded twice.
//
// Barrier protocol (L = leader CTA, rank 0; P = peer, rank 1):
//   full[s], b_full  live in L: count 2 = L's arrive.expect_tx(bytes of BOTH CTAs) + P's remote arrive; the TMA loads
//                    of both CTAs (cp.async.bulk.tensor ... .cta_group::2) credit L's barrier;
//   empty[s], acc_full[a]  live in both CTAs, signalled together by L's multicast tcgen05.commit;
//   acc_empty[a]     lives in L: count 2 x epilogue warps (P's warps arrive remotely).
#pragma once
#include "conv_igemm_v2.cuh"

namespace scm {

// CK = 64 only.  G.n_cta is the N HALF held by each CTA (G.n_total = 2 * G.n_cta); a "tile" is 256 plane rows.
template <int CK, int TPG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kV2Threads, 1)
conv3x3_igemm_v3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ IgemmParams P, const __grid_constant__ IgemmV2Geom G) {
    using Cfg = IgemmCfg<CK>;
    constexpr int RB = Cfg::kRowBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int chunks = P.cin_chunks;
    const int b_res_bytes = G.b_streamed ? 0 : 9 * chunks * G.b_tile_bytes;
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
    const uint32_t rank = cluster_ctarank();   // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1;
    const int n_full = 2 * G.n_cta;
    constexpr int kGroups = 9 / TPG;
    constexpr int kEpiArrivals = 2 * kV2EpiWarps;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < G.num_stages; ++s) {
            mbar_init(&full_bar[s], 2);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kEpiArrivals);
        }
        mbar_init(b_full, 2);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, 512);
        tmem_relinquish_pair();
    }
    // Programmatic dependent launch: everything up to pdl_sync() overlaps the previous kernel's tail.  With
    // P.prewait_weights the resident weight operand is part of that: it was packed at least two launches ago, and a
    // kernel can only start once its predecessor's CTAs have passed their own griddepcontrol.wait, i.e. once the
    // launch before THAT has completed - so the 147 KB weight load of every CTA runs while the chip still drains the
    // predecessor's last wave instead of in front of the first activation tile.
    const bool early_b = P.prewait_weights && !G.b_streamed;
    if (!early_b) pdl_sync();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // leader-side barrier addresses as seen from this CTA
    const uint32_t b_full_L = mapa_shared(smem_u32(b_full), 0);
    if (warp == 0) {
        if (!G.b_streamed && elect_one()) {
            // resident weights: this CTA's N half of every [tap][chunk] tile
            const uint32_t bytes = uint32_t(9 * chunks * G.n_cta * RB);
            if (rank == 0) mbar_arrive_expect_tx(b_full, 2 * bytes); else mbar_arrive_remote(b_full_L);
            for (int tap = 0; tap < 9; ++tap)
                for (int c = 0; c < chunks; ++c)
                    tma_load_2d_pair(s_b + size_t(tap * chunks + c) * G.b_tile_bytes, &tmap_b, b_full_L, c * CK,
                                     tap * n_full + int(rank) * G.n_cta);
        }
        __syncwarp();
    }
    if (early_b) pdl_sync();
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < n_full; i += 32 * kV2EpiWarps) s_bias[i] = (P.bias && i < P.bias_n) ? P.bias[i] : 0.f;
    }
    __syncthreads();  // s_bias

    if (warp == 0) {
        // ------------------------------ TMA producer (both CTAs) ------------------------------
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = uint32_t(G.loads * G.box_rows * RB) + (G.b_streamed ? 9u * uint32_t(G.n_cta * RB) : 0u);
        for (int tile = pair; tile < P.num_tiles; tile += G.tiles_stride) {
            const int m0 = tile * 256 + int(rank) * 128;
            for (int g = 0; g < kGroups; ++g) {
                const int tap0 = g * TPG;
                const int row0 = m0 + (tap0 / 3 - 1) * P.Wp + (tap0 % 3 - 1);
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        uint8_t* sa = s_a + size_t(stage) * G.a_stage_bytes;
                        const uint32_t full_L = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (P.debug & 32) {  // profiling: barrier traffic only, no TMA
                            if (rank == 0) mbar_arrive(&full_bar[stage]); else mbar_arrive_remote(full_L);
                        } else {
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * tx); else mbar_arrive_remote(full_L);
                            for (int l = 0; l < G.loads; ++l)
                                tma_load_2d_pair(sa + G.ld_smem[l], &tmap_a, full_L, P.a_c_off + c * CK,
                                                 row0 + G.ld_row[l]);
                            if (G.b_streamed) {  // Cin too large for resident weights: this chunk's nine tiles
                                for (int t = 0; t < 9; ++t)
                                    tma_load_2d_pair(sa + G.a_part_bytes + t * G.b_tile_bytes, &tmap_b, full_L, c * CK,
                                                     t * n_full + int(rank) * G.n_cta);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == G.num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer (leader CTA only) ------------------------------
        if (rank == 0) {
            const uint32_t idesc = make_idesc_ab(256, n_full, P.a_fmt, P.b_fmt, 0, 0);
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
            if (!G.b_streamed) mbar_wait(b_full, 0);
            const uint32_t a_part16 = uint32_t(G.a_part_bytes) >> 4;
            const uint32_t b_tap16 = G.b_streamed ? b_tile16 : uint32_t(chunks) * b_tile16;
            for (int tile = pair; tile < P.num_tiles; tile += G.tiles_stride) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * kAccStageCols);
                uint32_t accumulate = 0;
#pragma unroll 1
                for (int g = 0; g < kGroups; ++g) {
#pragma unroll 1
                    for (int c = 0; c < chunks; ++c) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t a_st = adesc0 + uint64_t(uint32_t(stage) * a_stage16);
                        const uint64_t b_st = G.b_streamed
                                                  ? a_st + uint64_t(a_part16)
                                                  : bdesc0 + uint64_t(uint32_t(g * TPG * chunks + c) * b_tile16);
                        if (elect_one()) {
#pragma unroll
                            for (int t = 0; t < TPG; ++t) {
                                const uint64_t at = a_st + uint64_t(a_off[t]);
                                const uint64_t bt = b_st + uint64_t(uint32_t(t) * b_tap16);
                                if (P.debug & 16) continue;  // profiling: no MMAs, commits only
#pragma unroll
                                for (int k = 0; k < Cfg::kKSteps; ++k)
                                    umma_f16_pair(tmem_d, at + uint64_t(2 * k), bt + uint64_t(2 * k), idesc,
                                                  (t == 0 && k == 0) ? accumulate : 1u);
                            }
                            umma_commit_pair(&empty_bar[stage], 3);
                        }
                        __syncwarp();
                        accumulate = 1;
                        if (++stage == G.num_stages) { stage = 0; phase ^= 1; }
                    }
                }
                if (elect_one()) umma_commit_pair(&acc_full[acc], 3);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue (igemm_epilogue.cuh), both CTAs ------------------------------
        const int q = warp & 3;
        const int ew = warp - 2;
        const int half = ew >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t acc_empty_L0 = mapa_shared(smem_u32(&acc_empty[0]), 0);
        const uint32_t acc_empty_L1 = mapa_shared(smem_u32(&acc_empty[1]), 0);
        for (int tile = pair; tile < P.num_tiles; tile += G.tiles_stride) {
            const int p = tile * 256 + int(rank) * 128 + q * 32 + lane;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kAccStageCols);
            igemm_epilogue_tile<32, false>(P, n_full, 0, n_full, p, half, lane, taddr, s_bias, &acc_full[acc],
                                           acc_phase, p + G.tiles_stride * 256);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(acc ? acc_empty_L1 : acc_empty_L0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace scm
