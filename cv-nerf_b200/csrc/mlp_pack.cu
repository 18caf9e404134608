// Parameter packing for the fused field-network kernel, and the hoisted view-direction term.
//
// nerf_pack_model: the 24 fp32 nn.Linear tensors of one Model (/root/reference/model.py:57-71)
//   -> BF16 stage images in UMMA K-major SWIZZLE_128B layout + fp32 tail (see mlp_layout.h).
// nerf_viewdir_term: l10's view-direction columns applied to PE4(viewdir) once per ray
//   (/root/reference/model.py:103-104 computes the same product once per *sample*).
#include <cuda_bf16.h>

#include "common.cuh"
#include "mlp_bwd_layout.h"
#include "mlp_layout.h"

namespace {

using namespace nerf;

struct ParamPtrs {
    const float* w[12];  // l1..l9, l_alpha, l10, l11
    const float* b[12];
};

// (stage) -> (layer, chunk, half)
__device__ __forceinline__ void stage_coords(int stage, int& l, int& chunk, int& half) {
    l = 0;
    int first = 0;
    while (l + 1 < kNumMmaLayers && first + layer_chunks(l) * layer_halves(l) <= stage) {
        first += layer_chunks(l) * layer_halves(l);
        ++l;
    }
    int local = stage - first;
    chunk = local / layer_halves(l);
    half = local % layer_halves(l);
}

// Source column of (layer l, K chunk, column k in chunk) in the nn.Linear weight, or -1 (zero pad).
__device__ __forceinline__ int source_col(int l, int chunk, int k) {
    if (l == 0) return k < kPeDim ? k : -1;                       // l1: [256,63]
    if (l == 5) return chunk == 0 ? (k < kPeDim ? k : -1)         // l6: PE columns first
                                  : kPeDim + (chunk - 1) * 64 + k;  //     then h5 columns 63..318
    return chunk * 64 + k;                                        // square layers, l10[:, :256]
}

__device__ __forceinline__ int param_index(int l) {  // MMA layer -> index in ParamPtrs (L1..L8 -> 0..7)
    return l;
}

__device__ __forceinline__ int in_features(int l) {
    return l == 0 ? 63 : (l == 5 ? 319 : 256);
}

// The folded layer (mlp_layout.h): W'[o][i..i+7] = sum_k l10.weight[o][k] * l9.weight[k][i..i+7], k = 0..255,
// FP32, ascending k.  Both blobs (forward stages and the transposed stages of the dZ chain) take their
// BF16 values from this one function, so they round the same FP32 numbers.
__device__ __forceinline__ void fold_row8(const ParamPtrs& p, int o, int i, float (&acc)[8]) {
    const float* w10 = p.w[10] + (size_t)o * 283;
    const float* w9 = p.w[8] + i;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // (unrolled so that the loads of eight k -- L2 hits right after an optimizer step -- are in flight together;
    // the accumulation order per output stays k = 0, 1, 2, ...)
#pragma unroll 8
    for (int k = 0; k < kHidden; ++k) {
        const float a = __ldg(w10 + k);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(w9 + (size_t)k * kHidden));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(w9 + (size_t)k * kHidden) + 1);
        acc[0] = fmaf(a, b0.x, acc[0]); acc[1] = fmaf(a, b0.y, acc[1]); acc[2] = fmaf(a, b0.z, acc[2]); acc[3] = fmaf(a, b0.w, acc[3]);
        acc[4] = fmaf(a, b1.x, acc[4]); acc[5] = fmaf(a, b1.y, acc[5]); acc[6] = fmaf(a, b1.z, acc[6]); acc[7] = fmaf(a, b1.w, acc[7]);
    }
}
// the same numbers for one input feature i and eight output rows o..o+7 (transposed stages)
__device__ __forceinline__ void fold_col8(const ParamPtrs& p, int o, int i, float (&acc)[8]) {
    const float* w10 = p.w[10] + (size_t)o * 283;
    const float* w9 = p.w[8] + i;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 8
    for (int k = 0; k < kHidden; ++k) {
        const float b = __ldg(w9 + (size_t)k * kHidden);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(__ldg(w10 + (size_t)e * 283 + k), b, acc[e]);
    }
}

// one thread = one 16-byte chunk (8 consecutive k) of one stage row
// fold_sel: 0 = every stage; 1 = skip the folded layer's stages; 2 = only them (gid counts from their first
// item).  The folded layer costs a 256-long dot product per value, so the one-launch re-pack of the training
// step gives those items blocks of their own, one warp each, spread over the SMs (eight warps of them in one
// block serialise on that SM's L1 port: 85 us instead of ~10).
constexpr int kFoldFwdFirst = layer_first_stage(kNumMmaLayers - 1) * kStageRows * 8;
constexpr int kFoldFwdItems = 4 * kStageRows * 8;
constexpr int kFoldBwdItems = 4 * kStageRows * 8;      // the transposed blob starts with them

__device__ __forceinline__ void pack_weights_item(const ParamPtrs& p, uint8_t* __restrict__ blob, int gid, int fold_sel = 0) {
    if (fold_sel == 2) gid += kFoldFwdFirst;
    if (gid >= kNumStages * kStageRows * 8) return;
    if (fold_sel == 1 && gid >= kFoldFwdFirst) return;
    int stage = gid / (kStageRows * 8);
    int r = (gid / 8) % kStageRows;
    int c16 = gid % 8;
    int l, chunk, half;
    stage_coords(stage, l, chunk, half);
    const int out_row = half * kStageRows + r;
    __nv_bfloat16 v[8];
    if (l == kNumMmaLayers - 1) {                 // l10' = l10[:, :256] . l9
        float acc[8];
        fold_row8(p, out_row, chunk * 64 + c16 * 8, acc);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(acc[e]);
    } else {
        const float* W = p.w[param_index(l)];
        const int ld = in_features(l);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int col = source_col(l, chunk, c16 * 8 + e);
            float x = col >= 0 ? W[(size_t)out_row * ld + col] : 0.f;
            v[e] = __float2bfloat16_rn(x);
        }
    }
    uint4 q;
    q.x = (uint32_t)__bfloat16_as_ushort(v[0]) | ((uint32_t)__bfloat16_as_ushort(v[1]) << 16);
    q.y = (uint32_t)__bfloat16_as_ushort(v[2]) | ((uint32_t)__bfloat16_as_ushort(v[3]) << 16);
    q.z = (uint32_t)__bfloat16_as_ushort(v[4]) | ((uint32_t)__bfloat16_as_ushort(v[5]) << 16);
    q.w = (uint32_t)__bfloat16_as_ushort(v[6]) | ((uint32_t)__bfloat16_as_ushort(v[7]) << 16);
    *reinterpret_cast<uint4*>(blob + (size_t)stage * kStageBytes + swz128_offset(r, c16 * 8)) = q;
}

// Transposed weights for the dZ chain (mlp_bwd_layout.h): stage row = input feature n, stage
// column = output feature k, value W[k][col_off + n].
__global__ void pack_weights_kernel(ParamPtrs p, uint8_t* __restrict__ blob) {
    pack_weights_item(p, blob, blockIdx.x * blockDim.x + threadIdx.x);
}

__device__ __forceinline__ void pack_weights_bwd_item(const ParamPtrs& p, uint8_t* __restrict__ blob, int gid, int fold_sel = 0) {
    if (gid >= kBwdStages * kStageRows * 8) return;
    if ((fold_sel == 1 && gid < kFoldBwdItems) || (fold_sel == 2 && gid >= kFoldBwdItems)) return;
    const int stage = gid / (kStageRows * 8);
    const int r = (gid / 8) % kStageRows;
    const int c16 = gid % 8;
    const int j = stage < 4 ? 0 : 1 + (stage - 4) / 8;
    const int local = stage - bwd_first_stage(j);
    const int chunk = local / 2, half = local % 2;
    const int n = half * kStageRows + r;
    uint32_t q[4];
    if (j == 0) {                                       // (l10[:, :256] . l9)^T: row = input feature of l9, column = output of l10
        float acc[8];
        fold_col8(p, chunk * 64 + c16 * 8, n, acc);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
            q[e] = *reinterpret_cast<uint32_t*>(&h);
        }
    } else {
        const int pi = 8 - j;                           // l8, l7, l6, l5, l4, l3, l2
        const int ld = pi == 5 ? 319 : 256;
        const int col_off = pi == 5 ? kPeDim : 0;       // l6: the h5 columns follow the 63 PE columns
        const float* W = p.w[pi];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = chunk * 64 + c16 * 8 + 2 * e;
            __nv_bfloat162 h = __floats2bfloat162_rn(W[(size_t)k * ld + col_off + n], W[(size_t)(k + 1) * ld + col_off + n]);
            q[e] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
    *reinterpret_cast<uint4*>(blob + (size_t)stage * kStageBytes + swz128_offset(r, c16 * 8)) =
        make_uint4(q[0], q[1], q[2], q[3]);
}

__global__ void pack_weights_bwd_kernel(ParamPtrs p, uint8_t* __restrict__ blob) {
    pack_weights_bwd_item(p, blob, blockIdx.x * blockDim.x + threadIdx.x);
}

__device__ __forceinline__ void pack_tail_bwd_item(const ParamPtrs& p, float* __restrict__ tail, int i) {
    if (i >= kBwdTailFloats) return;
    tail[i] = i < kBwdTailW11 ? p.w[9][i] : p.w[11][i - kBwdTailW11];
}

__global__ void pack_tail_bwd_kernel(ParamPtrs p, float* __restrict__ tail) {
    pack_tail_bwd_item(p, tail, blockIdx.x * blockDim.x + threadIdx.x);
}

// b10' = l10.bias + l10.weight[:, :256] . l9.bias for one output, by one warp (lanes stride over k)
__device__ __forceinline__ void pack_b10_warp(const ParamPtrs& p, float* __restrict__ tail, int o, int lane) {
    float v = 0.f;
#pragma unroll
    for (int k = lane; k < kHidden; k += 32) v = fmaf(__ldg(p.w[10] + (size_t)o * 283 + k), __ldg(p.b[8] + k), v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) tail[kTailB10 + o] = v + p.b[10][o];
}

__device__ __forceinline__ void pack_tail_item(const ParamPtrs& p, float* __restrict__ tail, int i, bool skip_b10 = false) {
    if (i >= kTailFloats) return;
    if (skip_b10 && i >= kTailB10) return;
    float v = 0.f;
    if (i < kTailWAlpha) {                       // biases b1..b8 (row 8 unused)
        v = i / kHidden < 8 ? p.b[i / kHidden][i % kHidden] : 0.f;
    } else if (i < kTailBAlpha) {
        v = p.w[9][i - kTailWAlpha];
    } else if (i < kTailW11) {
        v = (i == kTailBAlpha) ? p.b[9][0] : 0.f;
    } else if (i < kTailB11) {
        v = p.w[11][i - kTailW11];
    } else if (i < kTailW10View) {
        int k = i - kTailB11;
        v = k < 3 ? p.b[11][k] : 0.f;
    } else if (i < kTailB10) {
        int k = i - kTailW10View;
        int n = k / 28, c = k % 28;
        v = c < kViewPeDim ? p.w[10][(size_t)n * 283 + 256 + c] : 0.f;
    } else {                                     // l10.bias + l10.weight[:, :256] . l9.bias (l9 folded into l10)
        const int o = i - kTailB10;
        v = p.b[10][o];
#pragma unroll 16
        for (int k = 0; k < kHidden; ++k) v = fmaf(__ldg(p.w[10] + (size_t)o * 283 + k), __ldg(p.b[8] + k), v);
    }
    tail[i] = v;
}

// (the folded bias is always formed by pack_b10_warp, so that the stand-alone pack and the training step's
// one-launch re-pack write the same bits)
__global__ void pack_tail_kernel(ParamPtrs p, float* __restrict__ tail) {
    const int b10_blocks = kL10Out / 8;                      // eight warps per block, one output per warp
    if ((int)blockIdx.x < b10_blocks) {
        pack_b10_warp(p, tail, blockIdx.x * 8 + (threadIdx.x >> 5), threadIdx.x & 31);
        return;
    }
    pack_tail_item(p, tail, (blockIdx.x - b10_blocks) * blockDim.x + threadIdx.x, true);
}

// Everything the training step re-packs after an optimizer step, for up to two Models, in ONE
// launch: forward stages + tail and transposed stages + tail of each.
constexpr int kPackFwdBlocks = (kNumStages * kStageRows * 8 + 255) / 256;
constexpr int kPackBwdBlocks = (kBwdStages * kStageRows * 8 + 255) / 256;
constexpr int kPackTailBlocks = (kTailFloats + 255) / 256;
constexpr int kPackTailBwdBlocks = (kBwdTailFloats + 255) / 256;
constexpr int kPackFoldBlocks = (kFoldFwdItems + kFoldBwdItems) / 32 + kL10Out;   // one warp of fold items per block,
                                                                                   // then one warp per folded bias
constexpr int kPackBlocksPerModel = kPackFwdBlocks + kPackBwdBlocks + kPackTailBlocks + kPackTailBwdBlocks + kPackFoldBlocks;

struct PackAll {
    ParamPtrs p[2];
    uint8_t* fwd[2];
    uint8_t* bwd[2];
};

__global__ void __launch_bounds__(256) pack_all_kernel(const __grid_constant__ PackAll A) {
    const int model = blockIdx.x / kPackBlocksPerModel;
    int b = blockIdx.x % kPackBlocksPerModel;
    const ParamPtrs& p = A.p[model];
    if (b < kPackFoldBlocks) {              // first, so that the long items start first
        if (threadIdx.x >= 32) return;
        const int item = b * 32 + threadIdx.x;
        if (item < kFoldFwdItems) pack_weights_item(p, A.fwd[model], item, 2);
        else if (item < kFoldFwdItems + kFoldBwdItems) pack_weights_bwd_item(p, A.bwd[model], item - kFoldFwdItems, 2);
        else pack_b10_warp(p, reinterpret_cast<float*>(A.fwd[model] + kWeightBytes), b - (kFoldFwdItems + kFoldBwdItems) / 32, threadIdx.x);
        return;
    }
    b -= kPackFoldBlocks;
    if (b < kPackFwdBlocks) { pack_weights_item(p, A.fwd[model], b * 256 + threadIdx.x, 1); return; }
    b -= kPackFwdBlocks;
    if (b < kPackBwdBlocks) { pack_weights_bwd_item(p, A.bwd[model], b * 256 + threadIdx.x, 1); return; }
    b -= kPackBwdBlocks;
    if (b < kPackTailBlocks) {
        pack_tail_item(p, reinterpret_cast<float*>(A.fwd[model] + kWeightBytes), b * 256 + threadIdx.x, true);
        return;
    }
    b -= kPackTailBlocks;
    pack_tail_bwd_item(p, reinterpret_cast<float*>(A.bwd[model] + kBwdWeightBytes), b * 256 + threadIdx.x);
}

// block = 128 threads walks over kViewRows rows (rays) at a time; thread n produces out[row][n]
// with its 27 weights of W10[n, 256:283] held in registers for the whole block.
constexpr int kViewRows = 16;
__global__ void __launch_bounds__(128)
viewdir_term_kernel(const float* __restrict__ tail, const float* __restrict__ dirs, int dir_stride,
                    int embedded, long count, float* __restrict__ out) {
    __shared__ float pe[kViewRows][28];
    const int t = threadIdx.x;
    float w[kViewPeDim];
#pragma unroll
    for (int c = 0; c < kViewPeDim; ++c) w[c] = __ldg(tail + kTailW10View + t * 28 + c);
    const float b = __ldg(tail + kTailB10 + t);
    for (long row0 = (long)blockIdx.x * kViewRows; row0 < count; row0 += (long)gridDim.x * kViewRows) {
        __syncthreads();
        if (embedded) {
            for (int i = t; i < kViewRows * kViewPeDim; i += 128) {
                const int r = i / kViewPeDim, c = i % kViewPeDim;
                if (row0 + r < count) pe[r][c] = __ldg(dirs + (row0 + r) * dir_stride + c);
            }
        } else {
            // [x, sin(x 2^0), cos(x 2^0), ..., sin(x 2^3), cos(x 2^3)], model.py:15-31:
            // 16 rows x (3 raw + 12 sincos) work items over 128 threads
            for (int i = t; i < kViewRows * 15; i += 128) {
                const int r = i / 15, q = i % 15;
                if (row0 + r >= count) continue;
                const float* d = dirs + (row0 + r) * dir_stride;
                if (q < 3) {
                    pe[r][q] = __ldg(d + q);
                } else {
                    const int k = (q - 3) / 3, a = (q - 3) % 3;
                    float sn, cs;
                    sincosf(__ldg(d + a) * (float)(1 << k), &sn, &cs);
                    pe[r][3 + 6 * k + a] = sn;
                    pe[r][3 + 6 * k + 3 + a] = cs;
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < kViewRows; ++r) {
            if (row0 + r >= count) break;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < kViewPeDim; ++c) acc = fmaf(w[c], pe[r][c], acc);
            out[(row0 + r) * kL10Out + t] = acc + b;
        }
    }
}

// FreqEmbedding.embed as a standalone op: one thread per (row, input component).
__global__ void freq_encode_kernel(const float* __restrict__ x, long n, int dim, int n_freq,
                                   float* __restrict__ out) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * dim) return;
    long row = idx / dim;
    int a = (int)(idx % dim);
    float v = x[idx];
    float* o = out + row * (long)(dim * (1 + 2 * n_freq));
    o[a] = v;
    for (int k = 0; k < n_freq; ++k) {
        float s, c;
        sincosf(v * exp2f((float)k), &s, &c);
        o[dim + 2 * dim * k + a] = s;
        o[dim + 2 * dim * k + dim + a] = c;
    }
}

}  // namespace

extern "C" int nerf_freq_encode(const float* x, long n, int dim, int n_freq, float* out, void* stream) {
    nerf::DeviceGuard device_guard(out);
    if (n < 0 || dim < 1 || n_freq < 0 || (n > 0 && (!x || !out))) return nerf::arg_error("nerf_freq_encode");
    if (n == 0) return 0;
    freq_encode_kernel<<<nerf::blocks_for(n * dim, 256), 256, 0, (cudaStream_t)stream>>>(x, n, dim, n_freq, out);
    return nerf::check_launch("nerf_freq_encode");
}

extern "C" size_t nerf_packed_model_bytes(void) { return nerf::kPackedBytes; }

extern "C" int nerf_pack_model(const float* const* host_params, void* packed_out, void* stream) {
    nerf::DeviceGuard device_guard(packed_out);
    if (!host_params || !packed_out) return nerf::arg_error("nerf_pack_model");
    ParamPtrs p;
    for (int i = 0; i < 12; ++i) {
        p.w[i] = host_params[2 * i];
        p.b[i] = host_params[2 * i + 1];
        if (!p.w[i] || !p.b[i]) return nerf::arg_error("nerf_pack_model: null parameter");
    }
    cudaStream_t st = (cudaStream_t)stream;
    int n = kNumStages * kStageRows * 8;
    pack_weights_kernel<<<nerf::blocks_for(n, 256), 256, 0, st>>>(p, (uint8_t*)packed_out);
    pack_tail_kernel<<<nerf::blocks_for(kTailFloats, 256) + kL10Out / 8, 256, 0, st>>>(
        p, (float*)((uint8_t*)packed_out + kWeightBytes));
    return nerf::check_launch("nerf_pack_model");
}

extern "C" size_t nerf_packed_model_bwd_bytes(void) { return nerf::kBwdPackedBytes; }

extern "C" int nerf_pack_model_bwd(const float* const* host_params, void* packed_out, void* stream) {
    nerf::DeviceGuard device_guard(packed_out);
    if (!host_params || !packed_out) return nerf::arg_error("nerf_pack_model_bwd");
    ParamPtrs p;
    for (int i = 0; i < 12; ++i) {
        p.w[i] = host_params[2 * i];
        p.b[i] = host_params[2 * i + 1];
        if (!p.w[i] || !p.b[i]) return nerf::arg_error("nerf_pack_model_bwd: null parameter");
    }
    cudaStream_t st = (cudaStream_t)stream;
    int n = kBwdStages * kStageRows * 8;
    pack_weights_bwd_kernel<<<nerf::blocks_for(n, 256), 256, 0, st>>>(p, (uint8_t*)packed_out);
    pack_tail_bwd_kernel<<<nerf::blocks_for(kBwdTailFloats, 256), 256, 0, st>>>(
        p, (float*)((uint8_t*)packed_out + kBwdWeightBytes));
    return nerf::check_launch("nerf_pack_model_bwd");
}

extern "C" int nerf_pack_models_train(int n_models, const float* const* host_params, void* const* packed_out,
                                      void* const* packed_bwd_out, void* stream) {
    nerf::DeviceGuard device_guard((n_models > 0 && packed_out ? packed_out[0] : nullptr));
    if (n_models < 1 || n_models > 2 || !host_params || !packed_out || !packed_bwd_out)
        return nerf::arg_error("nerf_pack_models_train");
    PackAll A;
    for (int m = 0; m < n_models; ++m) {
        for (int i = 0; i < 12; ++i) {
            A.p[m].w[i] = host_params[m * 24 + 2 * i];
            A.p[m].b[i] = host_params[m * 24 + 2 * i + 1];
            if (!A.p[m].w[i] || !A.p[m].b[i]) return nerf::arg_error("nerf_pack_models_train: null parameter");
        }
        A.fwd[m] = (uint8_t*)packed_out[m];
        A.bwd[m] = (uint8_t*)packed_bwd_out[m];
        if (!A.fwd[m] || !A.bwd[m]) return nerf::arg_error("nerf_pack_models_train: null output");
    }
    pack_all_kernel<<<n_models * kPackBlocksPerModel, 256, 0, (cudaStream_t)stream>>>(A);
    return nerf::check_launch("nerf_pack_models_train");
}

extern "C" int nerf_viewdir_term(const void* packed, const float* dirs, int dir_stride, int embedded,
                                 long count, float* out, void* stream) {
    nerf::DeviceGuard device_guard(out);
    if (count < 0 || (count > 0 && (!packed || !dirs || !out))) return nerf::arg_error("nerf_viewdir_term");
    if (count == 0) return 0;
    const float* tail = (const float*)((const uint8_t*)packed + kWeightBytes);
    long blocks = (count + kViewRows - 1) / kViewRows;
    if (blocks > 148L * 16) blocks = 148L * 16;
    viewdir_term_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(tail, dirs, dir_stride, embedded,
                                                                           count, out);
    return nerf::check_launch("nerf_viewdir_term");
}
