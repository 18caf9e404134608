// Counter-based random draws for throughput runs (SURVEY.md App. A.7, section 7 hard part 5).
//
// The reference draws torch.rand / torch.randn tensors on the host side of every render_rays call
// (stratified jitter main.py:233, density noise main.py:188, inverse-CDF uniforms utils.py:23).
// Parity runs inject those tensors; throughput runs draw them inside the consuming kernel from
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) keyed by
// (seed; ray, group of 4 values, stream), so that no [n,S] tensor of random numbers is ever written
// to or read from HBM and a row-sharded render draws the same numbers as the unsharded one.
//
//   counter = (ray & 0xffffffff, ray >> 32, group, stream)      key = (seed & 0xffffffff, seed >> 32)
//   value i of a row lives in word (i & 3) of group (i >> 2)
//   uniform  = (word >> 8) * 2^-24                       in [0, 1)
//   normal   = Box-Muller on the word pairs (0,1) and (2,3): r = sqrt(-2 ln(((w0 >> 8) + 1) 2^-24)),
//              t = 2 pi (w1 >> 8) 2^-24, values (r cos t, r sin t)
#pragma once
#include <stdint.h>

#include "nerf_b200.h"

namespace nerf {

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), hi1 = __umulhi(0xCD9E8D57u, c.z);
#else
        const uint32_t hi0 = (uint32_t)((0xD2511F53ull * c.x) >> 32), hi1 = (uint32_t)((0xCD9E8D57ull * c.z) >> 32);
#endif
        const uint32_t lo0 = 0xD2511F53u * c.x, lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint4 draw_group(unsigned long long seed, int stream, long ray, int group) {
    return philox4x32_10(make_uint4((uint32_t)ray, (uint32_t)((unsigned long long)ray >> 32), (uint32_t)group,
                                    (uint32_t)stream),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ float to_uniform(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

__device__ __forceinline__ float4 uniform4(uint4 w) {
    return make_float4(to_uniform(w.x), to_uniform(w.y), to_uniform(w.z), to_uniform(w.w));
}

__device__ __forceinline__ float2 box_muller(uint32_t w0, uint32_t w1) {
    const float u1 = (float)((w0 >> 8) + 1u) * 5.9604644775390625e-8f;      // (0, 1]
    const float r = sqrtf(-2.f * logf(u1));
    float s, c;
    sincosf(6.283185307179586f * to_uniform(w1), &s, &c);
    return make_float2(r * c, r * s);
}

__device__ __forceinline__ float4 normal4(uint4 w) {
    const float2 a = box_muller(w.x, w.y), b = box_muller(w.z, w.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ float pick4(float4 v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

}  // namespace nerf
