// Backward of the field network, part 3: the CUDA-core pieces around the two tensor-core kernels.
//
//   nerf_mlp_bwd_heads    dW/db of l_alpha (256 -> 1) and l11 (128 -> 3): 640 MAC per sample
//   nerf_viewdir_term_bwd dW of l10's 27 view-direction columns and db of l10, through the per-ray
//                         sum of dZ10 (the transpose of the hoisting nerf_viewdir_term does)
//   nerf_grad_unpack      padded gradient blob -> the 24 .grad tensors of Model.parameters()
//   nerf_mse_loss_grad    mean((rgb - target)^2) and its gradient (/root/reference/main.py:380-383)
#include <cuda_bf16.h>

#include "common.cuh"
#include "mlp_bwd_layout.h"

namespace {

using namespace nerf;

__device__ __forceinline__ float bf16_at(const uint8_t* block0, int f, int r) {
    // element (row r, feature f) of a tile image whose 64-column blocks start at block0
    const uint8_t* p = block0 + (size_t)(f >> 6) * kBlockBytes + swz128_offset(r, f & 63);
    return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(p)) << 16);
}

// ------------------------------------------------------------------ l_alpha / l11
__global__ void __launch_bounds__(256) heads_bwd_kernel(const uint8_t* __restrict__ act,
                                                        const float* __restrict__ grad_raw, long M,
                                                        long n_tiles, float* __restrict__ grad) {
    __shared__ float4 g[kTileRows];
    const int f = threadIdx.x;
    float acc_a = 0.f, acc_r0 = 0.f, acc_r1 = 0.f, acc_r2 = 0.f, acc_b = 0.f;
    for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        __syncthreads();
        if (f < kTileRows) {
            const long row = t * kTileRows + f;
            g[f] = row < M ? __ldg(reinterpret_cast<const float4*>(grad_raw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const uint8_t* tile = act + (size_t)t * kActTileBytes;
        const uint8_t* h8 = tile + act_hidden(8);
        const uint8_t* h10 = tile + kActH10;
#pragma unroll 4
        for (int r = 0; r < kTileRows; ++r) {
            const float4 gr = g[r];
            acc_a = fmaf(gr.w, bf16_at(h8, f, r), acc_a);
            if (f < kL10Out) {
                const float h = bf16_at(h10, f, r);
                acc_r0 = fmaf(gr.x, h, acc_r0);
                acc_r1 = fmaf(gr.y, h, acc_r1);
                acc_r2 = fmaf(gr.z, h, acc_r2);
            }
        }
        if (f < 4) {
            for (int r = 0; r < kTileRows; ++r) acc_b += reinterpret_cast<const float*>(&g[r])[f];
        }
    }
    atomicAdd(grad + kG_WAlpha + f, acc_a);
    if (f < kL10Out) {
        atomicAdd(grad + kG_W11 + 0 * kL10Out + f, acc_r0);
        atomicAdd(grad + kG_W11 + 1 * kL10Out + f, acc_r1);
        atomicAdd(grad + kG_W11 + 2 * kL10Out + f, acc_r2);
    }
    if (f < 3) atomicAdd(grad + kG_B11 + f, acc_b);
    if (f == 3) atomicAdd(grad + kG_BAlpha, acc_b);
}

// ------------------------------------------------------------------ l10 view columns / bias
__global__ void __launch_bounds__(128) viewdir_term_bwd_kernel(const uint8_t* __restrict__ dz,
                                                               const float* __restrict__ dirs, int dir_stride,
                                                               int embedded, long M, int div, long count,
                                                               float* __restrict__ grad) {
    __shared__ float pe[28];
    const int j = threadIdx.x;
    float acc[kViewPeDim];
#pragma unroll
    for (int e = 0; e < kViewPeDim; ++e) acc[e] = 0.f;
    float accb = 0.f;
    for (long ray = blockIdx.x; ray < count; ray += gridDim.x) {
        __syncthreads();
        const float* d = dirs + ray * dir_stride;
        if (embedded) {
            if (j < kViewPeDim) pe[j] = __ldg(d + j);
        } else {
            if (j < 3) pe[j] = __ldg(d + j);
            if (j >= 32 && j < 32 + 12) {
                const int k = (j - 32) / 3, a = (j - 32) % 3;
                float s, c;
                sincosf(__ldg(d + a) * (float)(1 << k), &s, &c);
                pe[3 + 6 * k + a] = s;
                pe[3 + 6 * k + 3 + a] = c;
            }
        }
        __syncthreads();
        float dv = 0.f;
        const long r0 = ray * div, r1 = (r0 + div < M) ? r0 + div : M;
        for (long row = r0; row < r1; ++row) {
            const uint8_t* tile = dz + (size_t)(row >> 7) * kDzTileBytes + kDz10;
            dv += bf16_at(tile, j, (int)(row & 127));
        }
#pragma unroll
        for (int e = 0; e < kViewPeDim; ++e) acc[e] = fmaf(dv, pe[e], acc[e]);
        accb += dv;
    }
#pragma unroll
    for (int e = 0; e < kViewPeDim; ++e) atomicAdd(grad + kG_W10 + j * 288 + 256 + e, acc[e]);
    atomicAdd(grad + kG_B10 + j, accb);
}

// ------------------------------------------------------------------ blob -> .grad tensors
struct UnpackPtrs {
    float* p[NERF_N_PARAM_TENSORS];
};

struct Slot {
    int off, rows, cols, pitch, gap;   // gap: source column c >= gap is stored at c + 1 (l6's pad column)
};

__device__ __forceinline__ Slot grad_slot(int param) {
    const int layer = param >> 1;       // 0..8: l1..l9, 9: l_alpha, 10: l10, 11: l11
    const bool bias = param & 1;
    if (bias) {
        if (layer <= 8) return {kG_B + layer * 256, 1, 256, 256, 1 << 30};
        if (layer == 9) return {kG_BAlpha, 1, 1, 1, 1 << 30};
        if (layer == 10) return {kG_B10, 1, 128, 128, 1 << 30};
        return {kG_B11, 1, 3, 3, 1 << 30};
    }
    if (layer == 0) return {kG_W1, 256, 63, 64, 1 << 30};
    if (layer == 5) return {kG_W6, 256, 319, 320, 63};
    if (layer <= 8) return {grad_w_square(layer + 1), 256, 256, 256, 1 << 30};
    if (layer == 9) return {kG_WAlpha, 1, 256, 256, 1 << 30};
    if (layer == 10) return {kG_W10, 128, 283, 288, 1 << 30};
    return {kG_W11, 3, 128, 128, 1 << 30};
}

__global__ void grad_unpack_kernel(const float* __restrict__ blob, UnpackPtrs out, int accumulate) {
    const int param = blockIdx.y;
    const Slot s = grad_slot(param);
    const int n = s.rows * s.cols;
    float* dst = out.p[param];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / s.cols, c = i % s.cols;
        const float v = blob[s.off + r * s.pitch + (c >= s.gap ? c + 1 : c)];
        dst[i] = accumulate ? dst[i] + v : v;
    }
}

// ------------------------------------------------------------------ loss
__global__ void __launch_bounds__(256) mse_loss_grad_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ target, long n,
                                                            float inv_n, float* __restrict__ grad,
                                                            float* __restrict__ loss) {
    __shared__ float part[8];
    float acc = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float d = x[i] - target[i];
        acc = fmaf(d, d, acc);
        if (grad) grad[i] = 2.f * d * inv_n;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(loss, s * inv_n);
    }
}

}  // namespace

extern "C" int nerf_mlp_bwd_heads(const void* act_save, const float* grad_raw, long M, float* grad_blob,
                                  void* stream) {
    if (M < 0 || (M > 0 && (!act_save || !grad_raw || !grad_blob))) return nerf::arg_error("nerf_mlp_bwd_heads");
    if (M == 0) return 0;
    const long n_tiles = (M + kTileRows - 1) / kTileRows;
    int sms = nerf_b200_sm_count();
    if (sms <= 0) sms = 148;
    const long grid = n_tiles < 4L * sms ? n_tiles : 4L * sms;
    heads_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)act_save, grad_raw, M, n_tiles,
                                                                      grad_blob);
    return nerf::check_launch("nerf_mlp_bwd_heads");
}

extern "C" int nerf_viewdir_term_bwd(const void* dz, const float* dirs, int dir_stride, int embedded, long M,
                                     int vterm_div, float* grad_blob, void* stream) {
    if (M < 0 || vterm_div < 1 || (M > 0 && (!dz || !dirs || !grad_blob))) return nerf::arg_error("nerf_viewdir_term_bwd");
    if (M == 0) return 0;
    const long count = (M + vterm_div - 1) / vterm_div;
    int sms = nerf_b200_sm_count();
    if (sms <= 0) sms = 148;
    const long grid = count < 8L * sms ? count : 8L * sms;
    viewdir_term_bwd_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)dz, dirs, dir_stride,
                                                                             embedded, M, vterm_div, count, grad_blob);
    return nerf::check_launch("nerf_viewdir_term_bwd");
}

extern "C" int nerf_grad_unpack(const float* grad_blob, float* const* host_grads, int accumulate, void* stream) {
    if (!grad_blob || !host_grads) return nerf::arg_error("nerf_grad_unpack");
    UnpackPtrs out;
    for (int i = 0; i < NERF_N_PARAM_TENSORS; ++i) {
        out.p[i] = host_grads[i];
        if (!out.p[i]) return nerf::arg_error("nerf_grad_unpack: null gradient tensor");
    }
    grad_unpack_kernel<<<dim3(32, NERF_N_PARAM_TENSORS), 256, 0, (cudaStream_t)stream>>>(grad_blob, out, accumulate);
    return nerf::check_launch("nerf_grad_unpack");
}

extern "C" int nerf_mse_loss_grad(const float* x, const float* target, long n, float* grad_out, float* loss_accum,
                                  void* stream) {
    if (n < 0 || (n > 0 && (!x || !target))) return nerf::arg_error("nerf_mse_loss_grad");
    if (n == 0) return 0;
    long grid = (n + 255) / 256;
    if (grid > 1024) grid = 1024;
    mse_loss_grad_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, target, n, 1.f / (float)n, grad_out,
                                                                          loss_accum);
    return nerf::check_launch("nerf_mse_loss_grad");
}
