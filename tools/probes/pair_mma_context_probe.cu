// Probe: what slows tcgen05.mma.cta_group::2 down inside a kernel?  The pair MMA (M=256, N=256, K=16, SS form)
// issues at 128 cycles back to back (pair_mma_rate_probe); in the field kernel's context it ran at ~200.  This
// probe adds the kernel's companions one at a time (bit mask):
//   1  a multicast commit after every 4 MMAs          2  the accumulator alternates (columns 0 / 256) every 16 MMAs
//   4  the A operand walks over 8 blocks (128 KB)     8  16 warps per CTA loop tcgen05.ld.x16 + wait::ld on the accumulators
//   16 16 warps per CTA loop st.shared.v4 + fence.proxy.async into the A region
//   32 the commits of bit 1 go to the leader only (no multicast)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../cv-nerf_b200/csrc -I../../include -o pair_mma_context_probe pair_mma_context_probe.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <vector>
#include "umma.cuh"

constexpr int kIters = 64;
constexpr uint32_t kOffB = 131072;                      // A: 8 x 16 KB at 0, B: 4 x 16 KB (own halves)
constexpr uint32_t kOffBar = kOffB + 4 * 16384;
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 576;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) probe(int mode, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 64);
    volatile int* done = reinterpret_cast<volatile int*>(smem + kOffBar + 128);
    const uint32_t bar = sbase + kOffBar, bar_dummy = bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    for (uint32_t i = threadIdx.x; i < kOffBar / 4; i += kThreads)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (((i * 2654435761u) >> 20) & 0x00ff00ffu);
    if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::mbar_init(bar_dummy, 1); *done = 0; umma::fence_barrier_init(); }
    if (warp == 0) { umma::tmem_alloc_pair(umma::smem_u32(tmem_slot), 512); umma::tmem_relinquish_pair(); }
    umma::fence_proxy_async_smem();
    umma::tc_fence_before();
    umma::cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(256, 256);
        long long t0 = clock64();
        for (int it = 0; it < kIters; ++it) {
            const uint32_t d = tmem + ((mode & 2) ? (it & 1) * 256 : 0);
            const uint32_t a_base = sbase + ((mode & 4) ? (it & 1) * 65536 : 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma::mma_bf16_ss_pair(d, umma::smem_desc_sw128(a_base + j * 16384 + kk * 32),
                                           umma::smem_desc_sw128(sbase + kOffB + j * 16384 + kk * 32), idesc, (j | kk) ? 1u : 0u);
                if (mode & 32) umma::mma_commit_pair_local(bar_dummy);
                else if (mode & 1) umma::mma_commit_pair(bar_dummy);
            }
        }
        umma::mma_commit_pair(bar);
        long long t1 = clock64();
        umma::mbar_wait(bar, 0);
        long long t2 = clock64();
        cycles[(blockIdx.x >> 1) * 2] = t1 - t0;
        cycles[(blockIdx.x >> 1) * 2 + 1] = t2 - t0;
        *done = 1;
    } else if (warp == 1 && lane == 0) {
        umma::mbar_wait(bar, 0);      // the multicast commit also lands here
        *done = 1;
    } else if (warp >= 2) {
        const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        if (mode & 8) {
            while (!*done) {
                uint32_t v[16];
                umma::tmem_ld16(tacc + ((warp >> 2) & 15) * 16, v);
                umma::tmem_wait_ld();
                acc += v[0] ^ v[7] ^ v[15];
            }
        }
        if (mode & 16) {
            const uint32_t row_addr = sbase + ((warp - 2) >> 3) * 65536 + ((warp & 3) * 32 + lane) * 128;
            uint32_t k = 0;
            while (!*done) {
                umma::st_shared_v4(row_addr + (k & 3) * 16384 + ((k >> 2) & 7) * 16, k, k, k, k);
                ++k;
                if ((k & 15) == 0) umma::fence_proxy_async_smem();
            }
        }
        if (acc == 0x12345678u) cycles[0] = acc;
    }
    umma::tc_fence_before();
    umma::cluster_sync_all();
    if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc_pair(tmem, 512); }
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int grid = 148;
    for (int mode : {0, 1, 32, 2, 3, 4, 7, 8, 16, 24, 15, 31}) {
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, kThreads, kSmem>>>(mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s (mode %d)\n", cudaGetErrorString(e), mode); return 1; }
        }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double total = 0;
        for (int c = 0; c < grid / 2; ++c) total += h[c * 2 + 1];
        total /= grid / 2;
        const int n_mma = kIters * 16;
        printf("mode %2d: %.1f cycles/MMA\n", mode, total / n_mma);
    }
    return 0;
}
