// Probe: latencies of the hand-offs the field kernel's pipeline is built from (one CTA per SM, cycles).
//   1. mbarrier arrive -> waiter resumes, for a waiter that (a) spins on try_wait without a suspend hint,
//      (b) uses the 20 us suspend hint of umma::mbar_try_wait;
//   2. tcgen05.mma issue -> tcgen05.commit's arrive seen by a waiting thread, for 1, 4 and 16 MMAs
//      (M=128, N=256, K=16, SS form): pipeline fill + drain beyond the 128..160 cycles per MMA;
//   3. cp.async.bulk global(L2) -> shared: issue -> complete_tx seen, for 8, 16 and 32 KB, one copy in
//      flight, and the rate with 1, 2, 4, 8 copies of 16 KB in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../cv-nerf_b200/csrc -I../../include -o handoff_latency_probe handoff_latency_probe.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <vector>
#include "umma.cuh"

constexpr uint32_t kOffB = 65536;                 // A 4 x 16 KB | B 32 KB | fill 8 x 16 KB | barriers
constexpr uint32_t kOffFill = kOffB + 32768;
constexpr uint32_t kOffBar = kOffFill + 8 * 16384;
constexpr int kSmem = kOffBar + 512;
constexpr int kReps = 64;

__device__ __forceinline__ bool try_wait_nohint(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// out[0..1]: wake latency (no hint / hint); out[2..4]: commit latency for 1/4/16 MMAs; out[5..7]: bulk copy
// latency 8/16/32 KB; out[8..11]: bytes/cycle with 1/2/4/8 x 16 KB in flight
__global__ void __launch_bounds__(128) probe(const uint8_t* src, double* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t bar = sbase + kOffBar;            // [0] ping, [1] pong, [2] mma, [3..10] fill
    volatile long long* stamp = reinterpret_cast<volatile long long*>(smem + kOffBar + 256);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 384);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < kOffFill / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 11; ++i) umma::mbar_init(bar + 8 * i, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) { umma::tmem_alloc(umma::smem_u32(tmem_slot), 512); umma::tmem_relinquish(); }
    umma::fence_proxy_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- 1. arrive -> wake, ping-pong between warp 1 (arriver first) and warp 2 (waiter first) ----
    for (int mode = 0; mode < 2; ++mode) {
        long long sum = 0;
        if (warp == 1 && lane == 0) {
            for (int r = 0; r < kReps; ++r) {
                // wait a little so that the waiter is parked, then arrive with a time stamp
                long long t = clock64();
                while (clock64() - t < 3000) {}
                *stamp = clock64();
                __threadfence_block();
                umma::mbar_arrive(bar);
                // wait for the pong before the next round
                while (!try_wait_nohint(bar + 8, (uint32_t)((mode * kReps + r) & 1))) {}
            }
        } else if (warp == 2 && lane == 0) {
            for (int r = 0; r < kReps; ++r) {
                const uint32_t ph = (uint32_t)((mode * kReps + r) & 1);
                if (mode == 0) { while (!try_wait_nohint(bar, ph)) {} }
                else { while (!umma::mbar_try_wait(bar, ph)) {} }
                const long long t1 = clock64();
                sum += t1 - *stamp;
                umma::mbar_arrive(bar + 8);
            }
            if (blockIdx.x == 0) out[mode] = (double)sum / kReps;
        }
        __syncthreads();
    }

    // ---- 2. MMA issue -> commit observed ----
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(128, 256);
        uint32_t ph = 0;
        const int counts[3] = {1, 4, 16};
        for (int c = 0; c < 3; ++c) {
            long long sum = 0;
            for (int r = 0; r < kReps; ++r) {
                const long long t0 = clock64();
                for (int m = 0; m < counts[c]; ++m)
                    umma::mma_bf16_ss(tmem, umma::smem_desc_sw128(sbase + (m & 3) * 32), umma::smem_desc_sw128(sbase + kOffB + (m & 3) * 32),
                                      idesc, m ? 1u : 0u);
                umma::mma_commit(bar + 16);
                while (!try_wait_nohint(bar + 16, ph)) {}
                sum += clock64() - t0;
                ph ^= 1;
            }
            if (blockIdx.x == 0) out[2 + c] = (double)sum / kReps;
        }
    }
    __syncthreads();

    // ---- 3. bulk copy latency and rate ----
    if (warp == 1 && lane == 0) {
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t sizes[3] = {8192, 16384, 32768};
        for (int c = 0; c < 3; ++c) {
            long long sum = 0;
            for (int r = 0; r < kReps; ++r) {
                const long long t0 = clock64();
                umma::mbar_arrive_expect_tx(bar + 24, sizes[c]);
                umma::bulk_g2s(sbase + kOffFill, src + (size_t)((blockIdx.x * 5 + r) % 32) * 32768, sizes[c], bar + 24);
                while (!try_wait_nohint(bar + 24, ph[0])) {}
                sum += clock64() - t0;
                ph[0] ^= 1;
            }
            if (blockIdx.x == 0) out[5 + c] = (double)sum / kReps;
        }
        const int depths[4] = {1, 2, 4, 8};
        for (int c = 0; c < 4; ++c) {
            const int d = depths[c];
            const long long t0 = clock64();
            const int total = 256;
            for (int i = 0; i < total + d; ++i) {
                const int s = i % d;
                if (i >= d) { while (!try_wait_nohint(bar + 24 + 8 * s, ph[s])) {} ph[s] ^= 1; }
                if (i < total) {
                    umma::mbar_arrive_expect_tx(bar + 24 + 8 * s, 16384);
                    umma::bulk_g2s(sbase + kOffFill + s * 16384, src + (size_t)((blockIdx.x * 5 + i) % 64) * 16384, 16384, bar + 24 + 8 * s);
                }
            }
            if (blockIdx.x == 0) out[8 + c] = 16384.0 * total / (double)(clock64() - t0);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc(tmem, 512); }
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    uint8_t* src; double* out;
    cudaMalloc(&src, 32 * 32768); cudaMemset(src, 0x3c, 32 * 32768);
    cudaMalloc(&out, 16 * sizeof(double));
    for (int grid : {1, 148}) {
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, 128, kSmem>>>(src, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        }
        double h[16];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("grid %d (block 0):\n", grid);
        printf("  mbarrier arrive -> waiter resumes: spin try_wait %.0f cycles, try_wait with 20 us suspend hint %.0f cycles\n", h[0], h[1]);
        printf("  MMA issue -> commit observed: 1 MMA %.0f, 4 MMAs %.0f, 16 MMAs %.0f cycles (SS, N=256)\n", h[2], h[3], h[4]);
        printf("  cp.async.bulk L2 -> smem, one in flight: 8 KB %.0f, 16 KB %.0f, 32 KB %.0f cycles\n", h[5], h[6], h[7]);
        printf("  16 KB copies in flight 1/2/4/8: %.1f / %.1f / %.1f / %.1f B/cycle\n", h[8], h[9], h[10], h[11]);
    }
    return 0;
}
