// Probe: issue rate of tcgen05.mma (BF16, M=128, K=16) for the operand forms the field kernel can use,
// alone and beside the other users of the shared-memory / tensor-memory pipes:
//   form     0  SS N=256 (A and B from shared memory: production kernel)   1  SS N=128
//            2  TS N=128 (A from tensor memory)                            3  TS N=256
//            4  TS N=256, A slices 16 columns apart (the in-place layout of the TS field kernel)
//            5  ... and D / A regions swapped every 16 MMAs      6  ... and a commit every 4 MMAs
//   traffic  0  none
//            1  four warps drain D with tcgen05.ld (epilogue reads)
//            2  ... and write 16 packed columns per 32 read back with tcgen05.st (TS epilogue)
//            3  ... and (instead) write 64 B per 32 columns to shared memory (SS epilogue stores)
//            4  sixteen warps (4 per lane quadrant) run the TS epilogue pattern: ld 16 columns, st 8
//   fill     0  none     1  one thread streams 32 KB weight slots global(L2) -> shared with
//               cp.async.bulk, two in flight (the weight ring of the field kernel)
// One thread issues kIters x 16 MMAs back to back, commits once and waits; cycles per MMA are printed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../cv-nerf_b200/csrc -I../../include -o mma_rate_probe mma_rate_probe.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"

constexpr int kIters = 64;
constexpr uint32_t kOffB = 65536;                       // A: 4 x 16 KB at 0, B: 2 x 32 KB
constexpr uint32_t kOffFill = kOffB + 2 * 32768;        // 2 x 32 KB fill slots
constexpr uint32_t kOffBar = kOffFill + 2 * 32768;
constexpr int kSmem = kOffBar + 256;

__global__ void __launch_bounds__(576) probe(int form, int traffic, int fill, const uint8_t* src, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 64);
    volatile int* done = reinterpret_cast<volatile int*>(smem + kOffBar + 128);
    const uint32_t bar = sbase + kOffBar, bar_fill = bar + 8;   // bar_fill[2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < kOffBar / 4; i += 576)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (((i * 2654435761u) >> 20) & 0x00ff00ffu);
    if (threadIdx.x == 0) {
        *done = 0;
        umma::mbar_init(bar, 1);
        umma::mbar_init(bar_fill, 1);
        umma::mbar_init(bar_fill + 8, 1);
        umma::mbar_init(bar + 32, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) { umma::tmem_alloc(umma::smem_u32(tmem_slot), 512); umma::tmem_relinquish(); }
    umma::fence_proxy_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {   // fill the TMEM A region (columns 256..511) with finite values
        uint32_t w[8];
        for (int j = 0; j < 8; ++j) w[j] = 0x3c003c00u + threadIdx.x + j;
        for (int c = 256; c < 512; c += 8) umma::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + c, w);
        umma::tmem_wait_st();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    if (warp == 4) {
        if (lane == 0) {
            const int N = (form == 0 || form >= 3) ? 256 : 128;
            const uint32_t idesc = N == 256 ? umma::instr_desc_bf16(128, 256) : umma::instr_desc_bf16(128, 128);
            const bool ts = form >= 2;
            const uint32_t bar_dummy = bar + 32;
            long long t0 = clock64();
            for (int it = 0; it < kIters; ++it) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t b_desc = umma::smem_desc_sw128(sbase + kOffB + (j & 1) * 32768 + kk * 32);
                        const uint32_t acc = (j | kk) ? 1u : 0u;
                        if (form >= 4) {
                            const uint32_t flip = (form >= 5 && (it & 1)) ? 256u : 0u;
                            umma::mma_bf16_ts(tmem + flip, tmem + (256 ^ flip) + j * 64 + kk * 16, b_desc, idesc, acc);
                        } else if (ts) umma::mma_bf16_ts(tmem, tmem + 256 + j * 32 + kk * 8, b_desc, idesc, acc);
                        else umma::mma_bf16_ss(tmem, umma::smem_desc_sw128(sbase + j * 16384 + kk * 32), b_desc, idesc, acc);
                    }
                    if (form >= 6) umma::mma_commit(bar_dummy);
                }
            }
            umma::mma_commit(bar);      // one commit covers every MMA issued above
            long long t1 = clock64();
            umma::mbar_wait(bar, 0);
            long long t2 = clock64();
            cycles[blockIdx.x * 3] = t1 - t0;
            cycles[blockIdx.x * 3 + 1] = t2 - t0;
            *done = 1;
        }
    } else if (warp == 5) {
        if (lane == 0) {
            long long n = 0;
            if (fill) {
                for (int s = 0; s < 2; ++s) {
                    umma::mbar_arrive_expect_tx(bar_fill + 8 * s, 32768);
                    umma::bulk_g2s(sbase + kOffFill + s * 32768, src + ((blockIdx.x * 7 + s) % 36) * 32768, 32768, bar_fill + 8 * s);
                }
                while (!*done) {
                    const uint32_t s = (uint32_t)(n & 1), ph = (uint32_t)(n >> 1) & 1;
                    umma::mbar_wait(bar_fill + 8 * s, ph);
                    ++n;
                    umma::mbar_arrive_expect_tx(bar_fill + 8 * s, 32768);
                    umma::bulk_g2s(sbase + kOffFill + s * 32768, src + ((blockIdx.x * 7 + n + 1) % 36) * 32768, 32768, bar_fill + 8 * s);
                }
                // drain the two copies still in flight
                for (int k = 0; k < 2; ++k, ++n) umma::mbar_wait(bar_fill + 8 * (uint32_t)(n & 1), (uint32_t)(n >> 1) & 1);
            }
            cycles[blockIdx.x * 3 + 2] = n;
        }
    } else if (traffic == 4) {
        if (warp >= 6 || warp < 4) {
            const int w = warp < 4 ? warp : warp - 2;          // 16 crew warps: 0..15
            const int cg = w >> 2;
            const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            uint32_t sink = 0;
            while (!*done) {
                for (int j = 0; j < 4; ++j) {
                    uint32_t v[16], o[8];
                    umma::tmem_ld16(lane_base + j * 64 + cg * 16, v);
                    umma::tmem_wait_ld();
                    for (int e = 0; e < 8; ++e) o[e] = 0x3c003c00u | ((v[2 * e] ^ v[2 * e + 1]) & 0x00010001u);
                    sink ^= o[3];
                    umma::tmem_st8(lane_base + 256 + j * 64 + cg * 16, o);
                    umma::tmem_wait_st();
                }
            }
            if (sink == 0x12345u) cycles[0] = sink;
        }
    } else if (traffic >= 16) {
        // bit field: [0,2) crew warps = 4 << n; bit 2: 32-column loads; bit 3: + tcgen05.st of 8 columns;
        // bit 5: read the region the MMAs do NOT accumulate into (columns 256..511)
        const int n_w = 4 << (traffic & 3);
        const int w = warp < 4 ? warp : warp - 2;
        if ((warp >= 6 || warp < 4) && w < n_w) {
            const int cg = w >> 2;                                   // 0 .. n_w/4-1
            const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((traffic & 32) ? 256u : 0u);
            uint32_t sink = 0;
            long long loads = 0;
            while (!*done) {
                if (traffic & 4) {
                    for (int c = 0; c < 256; c += 128) {
                        uint32_t v[32];
                        umma::tmem_ld32(lane_base + c + cg * 32, v);
                        umma::tmem_wait_ld();
                        for (int e = 0; e < 32; ++e) sink ^= v[e];
                        loads += 32;
                    }
                } else {
                    for (int j = 0; j < 4; ++j) {
                        uint32_t v[16], o[8];
                        umma::tmem_ld16(lane_base + j * 64 + cg * 16, v);
                        umma::tmem_wait_ld();
                        for (int e = 0; e < 8; ++e) o[e] = 0x3c003c00u | ((v[2 * e] ^ v[2 * e + 1]) & 0x00010001u);
                        sink ^= o[3];
                        loads += 16;
                        if (traffic & 8) {
                            umma::tmem_st8((lane_base ^ 256u) + j * 64 + cg * 16, o);
                            umma::tmem_wait_st();
                        }
                    }
                }
            }
            if (sink == 0x12345u) cycles[0] = sink;
            if (lane == 0) atomicAdd((unsigned long long*)&cycles[148 * 3 + blockIdx.x], (unsigned long long)loads);
        }
    } else if (traffic && warp < 4) {
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t row_addr = sbase + (warp * 32 + lane) * 128;
        uint32_t sink = 0;
        while (!*done) {
            for (int c = 0; c < 256; c += 32) {
                uint32_t v[32];
                umma::tmem_ld32(lane_base + c, v);
                umma::tmem_wait_ld();
                for (int j = 0; j < 32; ++j) sink ^= v[j];
                if (traffic == 2) {
                    uint32_t o[8];
                    for (int j = 0; j < 8; ++j) o[j] = 0x3c003c00u | (v[j] & 0x00010001u);
                    umma::tmem_st8(lane_base + 256 + c / 2, o);
                    umma::tmem_st8(lane_base + 256 + c / 2 + 8, o);
                    umma::tmem_wait_st();
                } else if (traffic == 3) {
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t w = 0x3c003c00u | (v[q] & 0x00010001u);
                        umma::st_shared_v4(row_addr + (c >> 6) * 16384 + ((((c & 32) >> 1) + q * 16) ^ ((lane & 7) << 4)), w, w, w, w);
                    }
                }
            }
        }
        if (sink == 0x12345u) cycles[0] = sink;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc(tmem, 512); }
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    long long* d;
    cudaMalloc(&d, 148 * 4 * sizeof(long long));
    uint8_t* src;
    cudaMalloc(&src, 36 * 32768);
    cudaMemset(src, 0x3c, 36 * 32768);
    const char* names[7] = {"SS N=256", "SS N=128", "TS N=128", "TS N=256", "TS N=256 stride16", "TS N=256 stride16 flip", "TS N=256 stride16 flip commit4"};
    const int grid = 148;
    const int traffics[] = {0, 4, 16, 17, 18, 16 + 32, 17 + 32, 18 + 32, 17 + 4, 17 + 4 + 32, 18 + 8, 18 + 8 + 32};
    for (int fill = 1; fill < 2; ++fill)
        for (int traffic : traffics)
            for (int form : {0, 3, 4}) {
                cudaMemset(d, 0, 148 * 4 * sizeof(long long));
                for (int rep = 0; rep < 2; ++rep) {
                    probe<<<grid, 576, kSmem>>>(form, traffic, fill, src, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("form %d: CUDA error %s\n", form, cudaGetErrorString(e)); return 1; }
                }
                std::vector<long long> h(grid * 4);
                cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                double total = 0, copies = 0;
                for (int b = 0; b < grid; ++b) { total += h[b * 3 + 1]; copies += h[b * 3 + 2]; }
                total /= grid; copies /= grid;
                const int n_mma = kIters * 16;
                double lds = 0;
                for (int b = 0; b < grid; ++b) lds += h[grid * 3 + b];
                printf("fill %d traffic %2d  %-18s: %.1f cycles/MMA (%.0f FLOP/cycle/SM), fill %.1f B/cycle, tcgen05.ld %.1f B/cycle\n", fill, traffic,
                       names[form], total / n_mma, 2.0 * 128 * ((form == 0 || form >= 3) ? 256 : 128) * 16 * n_mma / total,
                       copies * 32768 / total, lds / grid / 2 * 128 * 4 / total);
            }
    return 0;
}
