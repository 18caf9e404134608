// Probe: layout of the A operand of tcgen05.mma when A lives in tensor memory (TS form).
// One CTA, M=128, N=32, K=16 BF16.  A[m][k] = small integers; written to TMEM by tcgen05.st.32x32b with
// the candidate layout "lane = row m, column = k/2, low half = even k"; B in shared memory (K-major,
// SWIZZLE_128B, 32 rows); D read back and compared on the host with A.B^T.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../cv-nerf_b200/csrc -I../../include -o ts_probe ts_mma_probe.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8192);
    const uint32_t bar = sbase + 8192 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = threadIdx.x;
    if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (warp == 0) { umma::tmem_alloc(umma::smem_u32(tmem_slot), 64); umma::tmem_relinquish(); }
    // B tile: [32 rows n][64 cols k] K-major swizzled image, only k < 16 used
    for (int i = threadIdx.x; i < 32 * 64; i += 128) {
        int n = i / 64, k = i % 64;
        __nv_bfloat16 v = k < 16 ? B[n * 16 + k] : __float2bfloat16(0.f);
        uint32_t off = n * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + ((k & 7) << 1);
        *reinterpret_cast<__nv_bfloat16*>(smem + off) = v;
    }
    umma::fence_proxy_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // A row -> 8 packed words
    uint32_t w[8];
    for (int j = 0; j < 8; ++j) {
        int k0 = variant == 0 ? 2 * j : j, k1 = variant == 0 ? 2 * j + 1 : j + 8;
        uint32_t lo = __bfloat16_as_ushort(A[row * 16 + k0]), hi = __bfloat16_as_ushort(A[row * 16 + k1]);
        w[j] = lo | (hi << 16);
    }
    const uint32_t a_addr = tmem + ((uint32_t)(warp * 32) << 16) + 32;   // columns 32..39
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a_addr),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(128, 32);
        const uint64_t b_desc = umma::smem_desc_sw128(sbase);
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem), "r"(tmem + 32), "l"(b_desc),
            "r"(idesc), "r"(0u) : "memory");
        umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();
    uint32_t v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    umma::tmem_wait_ld();
    for (int n = 0; n < 32; ++n) D[row * 32 + n] = __uint_as_float(v[n]);
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc(tmem, 64); }
    (void)lane;
}

int main() {
    std::vector<__nv_bfloat16> hA(128 * 16), hB(32 * 16);
    std::vector<float> fA(128 * 16), fB(32 * 16);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 17 - 8); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 13 - 6); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int variant = 0; variant < 2; ++variant) {
        probe<<<1, 128, 16384>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(128 * 32);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                float ref = 0;
                for (int k = 0; k < 16; ++k) ref += fA[m * 16 + k] * fB[n * 16 + k];
                double d = fabs(ref - hD[m * 32 + n]);
                if (d > worst) worst = d;
            }
        printf("variant %d (%s): max |D - A.B^T| = %g  D[0][0..3] = %g %g %g %g\n", variant,
               variant == 0 ? "word j = (k=2j, k=2j+1)" : "word j = (k=j, k=j+8)", worst, hD[0], hD[1], hD[2], hD[3]);
    }
    return 0;
}
