// Micro-benchmark: back-to-back tcgen05.mma issue rate on B200 (no TMA, operands are whatever is in smem).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate profiles/mma_rate.cu && /tmp/mma_rate
#include "../scm_gan_b200/csrc/ptx.cuh"
#include <cstdio>
using namespace scm;

template <int PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, int row_shift, long long* out, int mn_major = 0) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) { if (PAIR) { tmem_alloc_pair(&slot, 512); tmem_relinquish_pair(); } else { tmem_alloc(&slot, 512); tmem_relinquish(); } }
    tc_fence_before(); __syncthreads(); if (PAIR) cluster_sync_all(); tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, n, 1, mn_major, mn_major);
        // MN-major: two 64-channel atoms 16 KB apart (LBO), 8-pixel groups 1024 B apart (SBO)
        const uint64_t a0 = make_smem_desc(smem_u32(smem) + row_shift * 128, mn_major ? 16384 : 16, 1024, kLayoutSw128);
        const uint64_t b0 = make_smem_desc(smem_u32(smem) + 64 * 1024, mn_major ? 16384 : 16, 1024, kLayoutSw128);
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int adv = mn_major ? 128 * k : 2 * k;  // 16 pixel rows (2048 B) vs 32 B along K
                    if (PAIR) umma_f16_pair(tmem, a0 + adv, b0 + adv, idesc, 1);
                    else umma_f16(tmem, a0 + adv, b0 + adv, idesc, 1);
                }
            }
            if (PAIR) umma_commit_pair(&bar, 1); else umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads(); if (PAIR) cluster_sync_all();
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

int main() {
    long long* out; cudaMallocManaged(&out, 8);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int grid : {1, 148}) {
        for (int n : {16, 64, 128, 256}) {
            for (int shift : {0, 3}) {
                rate_kernel<0><<<grid, 128, smem>>>(n, iters, shift, out);
                cudaError_t e = cudaDeviceSynchronize();
                printf("1-CTA grid=%3d N=%3d shift=%d: %7.1f cycles/MMA (%s)\n", grid, n, shift, double(out[0]) / (4.0 * iters), cudaGetErrorString(e));
            }
        }
        for (int n : {16, 64, 128}) {
            for (int shift : {0, 1}) {
                rate_kernel<0><<<grid, 128, smem>>>(n, iters, shift, out, 1);
                cudaError_t e = cudaDeviceSynchronize();
                printf("1-CTA MN-major grid=%3d N=%3d shift=%d: %7.1f cycles/MMA (%s)\n", grid, n, shift, double(out[0]) / (4.0 * iters), cudaGetErrorString(e));
            }
        }
        for (int n : {128, 256}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid == 1 ? 2 : 148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {2, 1, 1};
            cfg.attrs = at; cfg.numAttrs = 1;
            int shift = 0;
            cudaLaunchKernelEx(&cfg, rate_kernel<1>, n, iters, shift, out, 0);
            cudaError_t e = cudaDeviceSynchronize();
            printf("2-CTA grid=%3d N=%3d (M=256): %7.1f cycles/MMA (%s)\n", cfg.gridDim.x, n, double(out[0]) / (4.0 * iters), cudaGetErrorString(e));
        }
    }
    return 0;
}
