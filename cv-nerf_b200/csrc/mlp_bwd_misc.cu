// Backward of the field network, part 3: the CUDA-core pieces around the two tensor-core kernels.
//
//   nerf_mlp_bwd_heads    dW/db of l_alpha (256 -> 1) and l11 (128 -> 3): 640 MAC per sample
//   nerf_viewdir_term_bwd dW of l10's 27 view-direction columns and db of l10, through the per-ray
//                         sum of dZ10 (the transpose of the hoisting nerf_viewdir_term does)
//   nerf_mlp_bwd_unfold   gradients of l9 and of l10's first 256 columns from G = dZ10^T h8 (l9 is
//                         folded into l10 in the forward and in the dZ chain, mlp_layout.h)
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

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xffff0000u);
    f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xffff0000u);
    f[4] = __uint_as_float(q.z << 16); f[5] = __uint_as_float(q.z & 0xffff0000u);
    f[6] = __uint_as_float(q.w << 16); f[7] = __uint_as_float(q.w & 0xffff0000u);
}

// ------------------------------------------------------------------ l_alpha / l11
// WARPS warps; warp w owns rows w, w+WARPS, ... of a tile, lane l owns the 16-byte chunk l of the row's
// h8 line (features 8l..8l+7) and, for l < 16, of its h10 line: whole 128-byte lines per request,
// 8 independent requests per lane in flight.  HBM-bound: 96 KB per tile.  WARPS = 4 (128 threads,
// <= 80 registers, 10 KB shared memory) fits on an SM BESIDE the persistent dZ-chain CTA, which leaves
// 10 K registers and ~30 KB: the training step launches it on a side stream under that tensor-bound
// kernel instead of next to the HBM-bound dW kernel.
constexpr int kHeadItems = 256 + 3 * 128 + 4;     // dW_alpha [256], dW11 [3][128], db11 [3], db_alpha
__device__ __forceinline__ int head_item_offset(int i) {
    return i < 256 ? kG_WAlpha + i : i < 256 + 384 ? kG_W11 + (i - 256) : (i - 640 < 3 ? kG_B11 + (i - 640) : kG_BAlpha);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) heads_bwd_kernel(const uint8_t* __restrict__ act,
                                                               const float* __restrict__ grad_raw, long M,
                                                               long n_tiles, float* __restrict__ grad,
                                                               float* __restrict__ partial) {
    __shared__ float red[WARPS][256 + 3 * 128 + 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a8[8], r0[8], r1[8], r2[8], gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 8; ++e) a8[e] = r0[e] = r1[e] = r2[e] = 0.f;
    const uint32_t blk_off = (uint32_t)(lane >> 3) * (uint32_t)kBlockBytes;
    const uint32_t c16 = lane & 7;
    for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint8_t* tile = act + (size_t)t * kActTileBytes;
        const uint8_t* h8 = tile + act_hidden(8) + blk_off;
        const uint8_t* h10 = tile + kActH10 + blk_off;
        // four rows per step: all twelve loads are issued before the first one is used
#pragma unroll 1
        for (int i = 0; i < kTileRows / WARPS; i += 4) {
            uint4 q8[4], q10[4];
            float4 g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = (i + u) * WARPS + warp;
                const long row = t * kTileRows + r;
                const uint32_t off = (uint32_t)r * 128 + ((c16 ^ (uint32_t)(r & 7)) << 4);
                q8[u] = ldg_nc_v4(h8 + off);
                q10[u] = lane < 16 ? ldg_nc_v4(h10 + off) : make_uint4(0u, 0u, 0u, 0u);
                g[u] = row < M ? __ldg(reinterpret_cast<const float4*>(grad_raw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float h[8];
                unpack8(q8[u], h);
#pragma unroll
                for (int e = 0; e < 8; ++e) a8[e] = fmaf(g[u].w, h[e], a8[e]);
                unpack8(q10[u], h);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    r0[e] = fmaf(g[u].x, h[e], r0[e]);
                    r1[e] = fmaf(g[u].y, h[e], r1[e]);
                    r2[e] = fmaf(g[u].z, h[e], r2[e]);
                }
                if (lane == 0) { gb[0] += g[u].x; gb[1] += g[u].y; gb[2] += g[u].z; gb[3] += g[u].w; }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[warp][lane * 8 + e] = a8[e];
        if (lane < 16) {
            red[warp][256 + 0 * 128 + lane * 8 + e] = r0[e];
            red[warp][256 + 1 * 128 + lane * 8 + e] = r1[e];
            red[warp][256 + 2 * 128 + lane * 8 + e] = r2[e];
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[warp][256 + 384 + k] = gb[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256 + 384 + 4; i += WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += red[w][i];
        if (partial) { partial[(size_t)blockIdx.x * kHeadItems + i] = s; continue; }     // deterministic mode
        atomicAdd(grad + head_item_offset(i), s);
    }
}

// Deterministic mode, second launches: blob element += its blocks' partial sums in block order.
__global__ void heads_reduce_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kHeadItems) return;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * kHeadItems + i];
    grad[head_item_offset(i)] += s;
}

// ------------------------------------------------------------------ l10 view columns / bias
// A block takes eight rays per step: each 16-lane group sums the dZ10 rows of ONE ray (16-byte loads,
// 128 features per row, eight independent loads in flight per thread), the per-ray sums are exchanged
// through shared memory, and thread j accumulates feature j's outer product with the rays' view
// encodings.  Two block barriers per eight rays.
__global__ void __launch_bounds__(128) viewdir_term_bwd_kernel(const uint8_t* __restrict__ dz,
                                                               const float* __restrict__ dirs, int dir_stride,
                                                               int embedded, long M, int div, long count,
                                                               float* __restrict__ grad, float* __restrict__ partial) {
    __shared__ float pe[8][28];
    __shared__ float part[8][128];
    const int j = threadIdx.x;
    const int sub = j >> 4;                 // ray slot of this 16-lane group
    const int l16 = j & 15;                 // chunk of the row: features 8*l16 .. 8*l16+7
    const uint32_t blk_off = (uint32_t)(l16 >> 3) * (uint32_t)kBlockBytes;
    const uint32_t c16 = l16 & 7;
    float acc[kViewPeDim];
#pragma unroll
    for (int e = 0; e < kViewPeDim; ++e) acc[e] = 0.f;
    float accb = 0.f;
    for (long ray0 = (long)blockIdx.x * 8; ray0 < count; ray0 += (long)gridDim.x * 8) {
        __syncthreads();                    // the previous step's reads of pe / part are done
        const long ray = ray0 + sub;
        const bool live = ray < count;
        const float* d = dirs + (live ? ray : 0) * dir_stride;
        if (embedded) {
            for (int e = l16; e < kViewPeDim; e += 16) pe[sub][e] = live ? __ldg(d + e) : 0.f;
        } else {
            if (l16 < 3) pe[sub][l16] = live ? __ldg(d + l16) : 0.f;
            if (l16 >= 3 && l16 < 15) {
                const int k = (l16 - 3) / 3, a = (l16 - 3) % 3;
                float sn, cs;
                sincosf((live ? __ldg(d + a) : 0.f) * (float)(1 << k), &sn, &cs);
                pe[sub][3 + 6 * k + a] = sn;
                pe[sub][3 + 6 * k + 3 + a] = cs;
            }
        }
        float s8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s8[e] = 0.f;
        if (live) {
            const long r0 = ray * div, r1 = (r0 + div < M) ? r0 + div : M;
#pragma unroll 8
            for (long row = r0; row < r1; ++row) {
                const int r = (int)(row & 127);
                const uint8_t* p = dz + (size_t)(row >> 7) * kDzTileBytes + kDz10 + blk_off + (uint32_t)r * 128 +
                                   ((c16 ^ (uint32_t)(r & 7)) << 4);
                float h[8];
                unpack8(ldg_nc_v4(p), h);
#pragma unroll
                for (int e = 0; e < 8; ++e) s8[e] += h[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) part[sub][l16 * 8 + e] = s8[e];
        __syncthreads();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float dv = part[g][j];
#pragma unroll
            for (int e = 0; e < kViewPeDim; ++e) acc[e] = fmaf(dv, pe[g][e], acc[e]);
            accb += dv;
        }
    }
    if (partial) {                          // deterministic mode: [block][128][28] = 27 view columns + bias
        float* pp = partial + ((size_t)blockIdx.x * 128 + j) * 28;
#pragma unroll
        for (int e = 0; e < kViewPeDim; ++e) pp[e] = acc[e];
        pp[kViewPeDim] = accb;
        return;
    }
#pragma unroll
    for (int e = 0; e < kViewPeDim; ++e) atomicAdd(grad + kG_W10 + j * 288 + 256 + e, acc[e]);
    atomicAdd(grad + kG_B10 + j, accb);
}

__global__ void view_reduce_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 128 * 28) return;
    const int j = i / 28, e = i % 28;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * 128 * 28 + i];
    grad[e < kViewPeDim ? kG_W10 + j * 288 + 256 + e : kG_B10 + j] += s;
}

// ------------------------------------------------------------------ unfold (l9 folded into l10)
// feat = W9 h8 + b9 feeds l10 without an activation (model.py:100-104), so with G = sum_rows dz10 (x) h8
// (accumulated by the dW kernel into the blob's scratch region) and db10 = sum_rows dz10:
//     dW10[:, :256] = G W9^T + db10 (x) b9          dW9 = W10a^T G          db9 = W10a^T db10
// FP32 on CUDA cores, 16.8 M multiply-adds per network and step.  One thread per output element; every
// element has a single writer, which ADDS to the blob like the other gradient kernels.
__global__ void __launch_bounds__(256) unfold_kernel(float* __restrict__ grad, const float* __restrict__ w9,
                                                     const float* __restrict__ b9, const float* __restrict__ w10) {
    const float* G = grad + kG_Fold;
    const float* db10 = grad + kG_B10;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < 128 * 256) {                                     // dW10[o][i], o < 128, i < 256
        const int o = t >> 8, i = t & 255;
        const float4* g4 = reinterpret_cast<const float4*>(G + o * 256);
        const float4* w4 = reinterpret_cast<const float4*>(w9 + (size_t)i * 256);
        float acc = db10[o] * __ldg(b9 + i);
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float4 g = g4[k], w = __ldg(w4 + k);
            acc = fmaf(g.x, w.x, acc); acc = fmaf(g.y, w.y, acc); acc = fmaf(g.z, w.z, acc); acc = fmaf(g.w, w.w, acc);
        }
        grad[kG_W10 + o * 288 + i] += acc;
    } else if (t < 128 * 256 + 256 * 256) {                  // dW9[i][k], i, k < 256
        const int u = t - 128 * 256;
        const int i = u >> 8, k = u & 255;
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < 128; ++o) acc = fmaf(__ldg(w10 + (size_t)o * 283 + i), G[o * 256 + k], acc);
        grad[grad_w_square(9) + i * 256 + k] += acc;
    } else if (t < 128 * 256 + 256 * 256 + 256) {            // db9[i]
        const int i = t - (128 * 256 + 256 * 256);
        float acc = 0.f;
        for (int o = 0; o < 128; ++o) acc = fmaf(__ldg(w10 + (size_t)o * 283 + i), db10[o], acc);
        grad[kG_B + 8 * 256 + i] += acc;
    }
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
                                                            float* __restrict__ loss, float* __restrict__ partial) {
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
        if (partial) partial[blockIdx.x] = s * inv_n;      // deterministic mode
        else atomicAdd(loss, s * inv_n);
    }
}

__global__ void loss_reduce_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ loss) {
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partial[b];
    *loss += s;
}

}  // namespace

static long heads_grid(long n_tiles) {
    int sms = nerf_b200_sm_count();
    if (sms <= 0) sms = 148;
    return n_tiles < 6L * sms ? n_tiles : 6L * sms;
}

static long view_grid(long count) {
    int sms = nerf_b200_sm_count();
    if (sms <= 0) sms = 148;
    const long groups = (count + 7) / 8;
    return groups < 8L * sms ? groups : 8L * sms;
}

static long loss_grid(long n) {
    const long grid = (n + 255) / 256;
    return grid > 1024 ? 1024 : grid;
}

static int launch_heads(const void* act_save, const float* grad_raw, long M, float* grad_blob, float* partial, void* stream) {
    nerf::DeviceGuard device_guard(grad_blob);
    if (M < 0 || (M > 0 && (!act_save || !grad_raw || !grad_blob))) return nerf::arg_error("nerf_mlp_bwd_heads");
    if (M == 0) return 0;
    const long n_tiles = (M + kTileRows - 1) / kTileRows;
    const long grid = heads_grid(n_tiles);
    heads_bwd_kernel<4><<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)act_save, grad_raw, M, n_tiles,
                                                                         grad_blob, partial);
    if (partial) heads_reduce_kernel<<<(kHeadItems + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial, (int)grid, grad_blob);
    return nerf::check_launch("nerf_mlp_bwd_heads");
}

static int launch_view(const void* dz, const float* dirs, int dir_stride, int embedded, long M, int vterm_div,
                       float* grad_blob, float* partial, void* stream) {
    nerf::DeviceGuard device_guard(grad_blob);
    if (M < 0 || vterm_div < 1 || (M > 0 && (!dz || !dirs || !grad_blob))) return nerf::arg_error("nerf_viewdir_term_bwd");
    if (M == 0) return 0;
    const long count = (M + vterm_div - 1) / vterm_div;
    const long grid = view_grid(count);
    viewdir_term_bwd_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)dz, dirs, dir_stride,
                                                                             embedded, M, vterm_div, count, grad_blob, partial);
    if (partial) view_reduce_kernel<<<(128 * 28 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial, (int)grid, grad_blob);
    return nerf::check_launch("nerf_viewdir_term_bwd");
}

extern "C" int nerf_mlp_bwd_heads(const void* act_save, const float* grad_raw, long M, float* grad_blob,
                                  void* stream) {
    return launch_heads(act_save, grad_raw, M, grad_blob, nullptr, stream);
}

extern "C" int nerf_viewdir_term_bwd(const void* dz, const float* dirs, int dir_stride, int embedded, long M,
                                     int vterm_div, float* grad_blob, void* stream) {
    return launch_view(dz, dirs, dir_stride, embedded, M, vterm_div, grad_blob, nullptr, stream);
}

// ---- deterministic accumulation (parity / debugging runs) ---------------------------------------------
// The kernels above add per-block sums to the blob with floating-point atomics, so two runs differ in the
// last bits of a gradient.  The _det variants write the per-block partial sums to `scratch` and a second
// launch adds them to the blob in block order: bitwise reproducible on a given GPU model.
// nerf_bwd_det_scratch_bytes(kind, size): kind 1 heads (size = M), 2 view columns (size = rays), 3 loss
// (size = n); the dW kernel has nerf_mlp_bwd_dw_det_scratch_bytes().
extern "C" size_t nerf_bwd_det_scratch_bytes(int kind, long size) {
    if (size <= 0) return 0;
    if (kind == 1) return (size_t)heads_grid((size + kTileRows - 1) / kTileRows) * kHeadItems * 4;
    if (kind == 2) return (size_t)view_grid(size) * 128 * 28 * 4;
    if (kind == 3) return (size_t)loss_grid(size) * 4;
    return 0;
}

extern "C" int nerf_mlp_bwd_heads_det(const void* act_save, const float* grad_raw, long M, float* grad_blob,
                                      void* scratch, void* stream) {
    if (!scratch) return nerf::arg_error("nerf_mlp_bwd_heads_det: scratch");
    return launch_heads(act_save, grad_raw, M, grad_blob, (float*)scratch, stream);
}

extern "C" int nerf_viewdir_term_bwd_det(const void* dz, const float* dirs, int dir_stride, int embedded, long M,
                                         int vterm_div, float* grad_blob, void* scratch, void* stream) {
    if (!scratch) return nerf::arg_error("nerf_viewdir_term_bwd_det: scratch");
    return launch_view(dz, dirs, dir_stride, embedded, M, vterm_div, grad_blob, (float*)scratch, stream);
}

extern "C" int nerf_mlp_bwd_unfold(float* grad_blob, const float* l9_weight, const float* l9_bias,
                                   const float* l10_weight, void* stream) {
    nerf::DeviceGuard device_guard(grad_blob);
    if (!grad_blob || !l9_weight || !l9_bias || !l10_weight) return nerf::arg_error("nerf_mlp_bwd_unfold");
    constexpr int kItems = 128 * 256 + 256 * 256 + 256;
    unfold_kernel<<<(kItems + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grad_blob, l9_weight, l9_bias, l10_weight);
    return nerf::check_launch("nerf_mlp_bwd_unfold");
}

extern "C" int nerf_grad_unpack(const float* grad_blob, float* const* host_grads, int accumulate, void* stream) {
    nerf::DeviceGuard device_guard(grad_blob);
    if (!grad_blob || !host_grads) return nerf::arg_error("nerf_grad_unpack");
    UnpackPtrs out;
    for (int i = 0; i < NERF_N_PARAM_TENSORS; ++i) {
        out.p[i] = host_grads[i];
        if (!out.p[i]) return nerf::arg_error("nerf_grad_unpack: null gradient tensor");
    }
    grad_unpack_kernel<<<dim3(32, NERF_N_PARAM_TENSORS), 256, 0, (cudaStream_t)stream>>>(grad_blob, out, accumulate);
    return nerf::check_launch("nerf_grad_unpack");
}

static int launch_loss(const float* x, const float* target, long n, float* grad_out, float* loss_accum, float* partial,
                       void* stream) {
    nerf::DeviceGuard device_guard(loss_accum);
    if (n < 0 || (n > 0 && (!x || !target))) return nerf::arg_error("nerf_mse_loss_grad");
    if (n == 0) return 0;
    const long grid = loss_grid(n);
    mse_loss_grad_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, target, n, 1.f / (float)n, grad_out,
                                                                          loss_accum, partial);
    if (partial && loss_accum) loss_reduce_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(partial, (int)grid, loss_accum);
    return nerf::check_launch("nerf_mse_loss_grad");
}

extern "C" int nerf_mse_loss_grad(const float* x, const float* target, long n, float* grad_out, float* loss_accum,
                                  void* stream) {
    return launch_loss(x, target, n, grad_out, loss_accum, nullptr, stream);
}

extern "C" int nerf_mse_loss_grad_det(const float* x, const float* target, long n, float* grad_out, float* loss_accum,
                                      void* scratch, void* stream) {
    if (!scratch) return nerf::arg_error("nerf_mse_loss_grad_det: scratch");
    return launch_loss(x, target, n, grad_out, loss_accum, (float*)scratch, stream);
}
