// Probe: tcgen05.mma (SS form, M=128, N=32, K=16) with the B operand in a K-major SWIZZLE_64B tile
// ([rows][32 bf16] = 64-byte rows, 16-byte chunk index XOR ((row >> 1) & 3), 8-row groups 512 B apart)
// while A stays in the SWIZZLE_128B [128][64] tile layout of the field kernel.  Checks D = A.B^T for
// both K = 16 slices of the 32-wide B tile (descriptor start address + 32 bytes for the second).
// Also: fill rate of 16 KB cp.async.bulk copies issued by ONE vs TWO threads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../cv-nerf_b200/csrc -I../../include -o sw64_mma_probe sw64_mma_probe.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"

__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;        // SBO: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    // A tile [128][64] SW128 at 0 (16 KB), B tile [32][32] SW64 at 16 KB (2 KB)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 20480);
    const uint32_t bar = sbase + 20480 + 16;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (warp == 0) { umma::tmem_alloc(umma::smem_u32(tmem_slot), 64); umma::tmem_relinquish(); }
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        int r = i / 64, k = i % 64;
        __nv_bfloat16 v = k < 32 ? A[r * 32 + k] : __float2bfloat16(0.f);
        uint32_t off = r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1);
        *reinterpret_cast<__nv_bfloat16*>(smem + off) = v;
    }
    for (int i = threadIdx.x; i < 32 * 32; i += 128) {
        int n = i / 32, k = i % 32;
        uint32_t off = n * 64 + ((((k >> 3) ^ ((n >> 1) & 3)) & 3) << 4) + ((k & 7) << 1);
        *reinterpret_cast<__nv_bfloat16*>(smem + 16384 + off) = B[n * 32 + k];
    }
    umma::fence_proxy_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(128, 32);
        for (int kk = 0; kk < 2; ++kk)
            umma::mma_bf16_ss(tmem, umma::smem_desc_sw128(sbase + kk * 32), smem_desc_sw64(sbase + 16384 + kk * 32), idesc, kk ? 1u : 0u);
        umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();
    uint32_t v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    umma::tmem_wait_ld();
    for (int n = 0; n < 32; ++n) D[threadIdx.x * 32 + n] = __uint_as_float(v[n]);
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc(tmem, 64); }
}

// fill rate: `nthr` threads (one per warp) each keep `depth` 16 KB copies in flight
__global__ void __launch_bounds__(128) fill_rate(const uint8_t* src, int nthr, int depth, double* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t bar = sbase + 8 * 16384;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) umma::mbar_init(bar + 8 * i, 1); umma::fence_barrier_init(); }
    __syncthreads();
    if (lane == 0 && warp < nthr) {
        uint32_t ph[4] = {0, 0, 0, 0};
        const int total = 256;
        const long long t0 = clock64();
        for (int i = 0; i < total + depth; ++i) {
            const int s = i % depth, slot = warp * 4 + s;
            if (i >= depth) { umma::mbar_wait(bar + 8 * slot, ph[s]); ph[s] ^= 1; }
            if (i < total) {
                umma::mbar_arrive_expect_tx(bar + 8 * slot, 16384);
                umma::bulk_g2s(sbase + slot * 16384, src + (size_t)((blockIdx.x * 5 + i * nthr + warp) % 64) * 16384, 16384, bar + 8 * slot);
            }
        }
        if (blockIdx.x == 0 && warp == 0) out[0] = 16384.0 * total * nthr / (double)(clock64() - t0);
    }
}

int main() {
    std::vector<__nv_bfloat16> hA(128 * 32), hB(32 * 32);
    std::vector<float> fA(128 * 32), fB(32 * 32);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 17 - 8); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 13 - 6); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576);
    probe<<<1, 128, 24576>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(128 * 32);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            float ref = 0;
            for (int k = 0; k < 32; ++k) ref += fA[m * 32 + k] * fB[n * 32 + k];
            worst = fmax(worst, fabs(ref - hD[m * 32 + n]));
        }
    printf("SW64 B tile, K = 32 in two MMAs: max |D - A.B^T| = %g  D[0][0..3] = %g %g %g %g\n", worst, hD[0], hD[1], hD[2], hD[3]);

    uint8_t* src; double* out;
    cudaMalloc(&src, 64 * 16384); cudaMemset(src, 1, 64 * 16384); cudaMalloc(&out, 8);
    cudaFuncSetAttribute(fill_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 256);
    for (int nthr = 1; nthr <= 2; ++nthr)
        for (int depth : {2, 4}) {
            for (int rep = 0; rep < 2; ++rep) fill_rate<<<148, 128, 8 * 16384 + 256>>>(src, nthr, depth, out);
            cudaDeviceSynchronize();
            double h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("16 KB copies, %d issuing thread(s) x %d in flight: %.1f B/cycle per SM\n", nthr, depth, h);
        }
    return 0;
}
