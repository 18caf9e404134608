// Probe: issue rate of tcgen05.mma.cta_group::2 (BF16, M=256 over a CTA pair, K=16) for N=256 and
// N=128, SS form, each CTA holding its half of B ([N/2 rows][64] K-major SWIZZLE_128B).  The leader
// issues kIters x 16 MMAs back to back, commits once (multicast) and waits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../cv-nerf_b200/csrc -I../../include -o pair_mma_rate_probe pair_mma_rate_probe.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <vector>
#include "umma.cuh"

constexpr int kIters = 64;
constexpr uint32_t kOffB = 65536;                       // A: 4 x 16 KB at 0, B: 4 x 16 KB (own halves)
constexpr uint32_t kOffBar = kOffB + 4 * 16384;
constexpr int kSmem = kOffBar + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe(int n256, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 64);
    const uint32_t bar = sbase + kOffBar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    for (uint32_t i = threadIdx.x; i < kOffBar / 4; i += 128)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (((i * 2654435761u) >> 20) & 0x00ff00ffu);
    if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (warp == 0) { umma::tmem_alloc_pair(umma::smem_u32(tmem_slot), 512); umma::tmem_relinquish_pair(); }
    umma::fence_proxy_async_smem();
    umma::tc_fence_before();
    umma::cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = n256 ? umma::instr_desc_bf16(256, 256) : umma::instr_desc_bf16(256, 128);
        long long t0 = clock64();
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma::mma_bf16_ss_pair(tmem, umma::smem_desc_sw128(sbase + j * 16384 + kk * 32),
                                           umma::smem_desc_sw128(sbase + kOffB + j * 16384 + kk * 32), idesc, (j | kk) ? 1u : 0u);
            }
        }
        umma::mma_commit_pair(bar);
        long long t1 = clock64();
        umma::mbar_wait(bar, 0);
        long long t2 = clock64();
        cycles[(blockIdx.x >> 1) * 2] = t1 - t0;
        cycles[(blockIdx.x >> 1) * 2 + 1] = t2 - t0;
    } else if (warp == 1 && lane == 0) {
        umma::mbar_wait(bar, 0);      // the multicast commit also lands here
    }
    umma::tc_fence_before();
    umma::cluster_sync_all();
    if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc_pair(tmem, 512); }
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    for (int grid : {2, 148}) {
        for (int n256 = 1; n256 >= 0; --n256) {
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<grid, 128, kSmem>>>(n256, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            }
            std::vector<long long> h(grid);
            cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double total = 0;
            for (int c = 0; c < grid / 2; ++c) total += h[c * 2 + 1];
            total /= grid / 2;
            const int n_mma = kIters * 16;
            printf("grid %3d  pair SS M=256 N=%d: %.1f cycles/MMA (%.0f FLOP/cycle/SM)\n", grid, n256 ? 256 : 128, total / n_mma,
                   2.0 * 128 * (n256 ? 256 : 128) * 16 * n_mma / total);
        }
    }
    return 0;
}
