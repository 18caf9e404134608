// Optimizer and train-loop front end.
//
//   nerf_adam_step        torch.optim.Adam(lr, betas, eps) of /root/reference/main.py:144,386 as ONE
//                         multi-tensor launch over all parameter tensors (the reference steps 48
//                         tensors with eager ops)
//   nerf_adam_step_blob   the same, reading the gradient straight from the padded gradient blob the
//                         backward kernels accumulate into (and the data-parallel all-reduce sums)
//   nerf_train_rays       the ray/target batch of one train iteration, main.py:351-374, generated on
//                         the device for the chosen pixels only (the reference builds every ray of the
//                         image, a full meshgrid and np.random.choice on the host each iteration)
#include "common.cuh"
#include "mlp_bwd_layout.h"

namespace {

using namespace nerf;

constexpr int kMaxTensors = 48;

struct AdamTensors {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* m[kMaxTensors];
    float* v[kMaxTensors];
    int n[kMaxTensors];
};

struct AdamScalars {
    float lr, beta1, beta2, eps;
    float bias_correction1;        // 1 - beta1^t
    float bias_correction2_sqrt;   // sqrt(1 - beta2^t)
    float grad_scale;              // 1 / world size for an all-reduced (summed) gradient
};

// torch/optim/adam.py _single_tensor_adam, op for op:
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(bc2) + eps;
//   p.addcdiv_(m, denom, value=-(lr/bc1))
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamScalars& s) {
    m = m + (1.f - s.beta1) * (g - m);
    v = v * s.beta2 + (1.f - s.beta2) * g * g;
    const float denom = sqrtf(v) / s.bias_correction2_sqrt + s.eps;
    p = p - (s.lr / s.bias_correction1) * (m / denom);
}

__global__ void __launch_bounds__(256) adam_tensors_kernel(const __grid_constant__ AdamTensors T,
                                                           const AdamScalars s) {
    const int t = blockIdx.y;
    const int n = T.n[t];
    float* __restrict__ p = T.p[t];
    const float* __restrict__ g = T.g[t];
    float* __restrict__ m = T.m[t];
    float* __restrict__ v = T.v[t];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_update(pi, g[i] * s.grad_scale, mi, vi, s);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

struct BlobTensors {
    float* p[NERF_N_PARAM_TENSORS];
    float* m[NERF_N_PARAM_TENSORS];
    float* v[NERF_N_PARAM_TENSORS];
};

struct Slot {
    int off, rows, cols, pitch, gap;
};

__device__ __forceinline__ Slot grad_slot(int param) {   // same table as nerf_grad_unpack
    const int layer = param >> 1;
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

__global__ void __launch_bounds__(256) adam_blob_kernel(const float* __restrict__ blob,
                                                        const __grid_constant__ BlobTensors T,
                                                        const AdamScalars s) {
    const int t = blockIdx.y;
    const Slot sl = grad_slot(t);
    const int n = sl.rows * sl.cols;
    float* __restrict__ p = T.p[t];
    float* __restrict__ m = T.m[t];
    float* __restrict__ v = T.v[t];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / sl.cols, c = i % sl.cols;
        const float g = blob[sl.off + r * sl.pitch + (c >= sl.gap ? c + 1 : c)] * s.grad_scale;
        float pi = p[i], mi = m[i], vi = v[i];
        adam_update(pi, g, mi, vi, s);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// Data-parallel variant fused with the gradient exchange: every rank's gradient blob lives in
// peer-mapped (symmetric) memory; each rank reads all `world` blobs straight over NVLink (plain
// ld.global on the mapped peer pointers), sums them in rank order -- the same order on every rank,
// so the replicas stay bit-identical -- and applies Adam to its own copy of the parameters.  One
// launch replaces ncclAllReduce + Adam; 4.8 MB x world cross NVLink per rank and step.
constexpr int kMaxPeers = 16;
struct PeerBlobs {
    const float* blob[kMaxPeers];
    int world;
};

__global__ void __launch_bounds__(256) adam_blob_peers_kernel(const __grid_constant__ PeerBlobs B,
                                                              const __grid_constant__ BlobTensors T,
                                                              const AdamScalars s) {
    const int t = blockIdx.y;
    const Slot sl = grad_slot(t);
    const int n = sl.rows * sl.cols;
    float* __restrict__ p = T.p[t];
    float* __restrict__ m = T.m[t];
    float* __restrict__ v = T.v[t];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / sl.cols, c = i % sl.cols;
        const int off = sl.off + r * sl.pitch + (c >= sl.gap ? c + 1 : c);
        float g = 0.f;
        for (int k = 0; k < B.world; ++k) g += __ldcv(B.blob[k] + off);   // peer memory: never cached stale
        g *= s.grad_scale;
        float pi = p[i], mi = m[i], vi = v[i];
        adam_update(pi, g, mi, vi, s);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// ------------------------------------------------------------------ train-loop ray batch
// Pseudo-random permutation of [0, 4^b) (4-round Feistel on b+b bits); cycle-walking restricts it
// to [0, domain): distinct inputs give distinct outputs, i.e. sampling WITHOUT replacement
// (np.random.choice(..., replace=False), main.py:368) with O(1) work and no memory per sample.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t feistel(uint32_t x, int half_bits, uint64_t key) {
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
        const uint32_t k = (uint32_t)(key >> (16 * round)) ^ (0x9e3779b9u * (round + 1));
        const uint32_t f = mix32(r ^ k) & mask;
        const uint32_t nl = r;
        r = l ^ f;
        l = nl;
    }
    return (l << half_bits) | r;
}

struct Pose {
    float r[3][3];
    float t[3];
};

__global__ void train_rays_kernel(int H, int W, float half_h, float half_w, float f, float cw, float ch,
                                  const float* __restrict__ pose, const int* __restrict__ pix_in,
                                  uint64_t key, int r0, int c0, int crop_h, int crop_w, int half_bits, long n,
                                  int ndc, float near, float far, const float* __restrict__ image,
                                  float* __restrict__ rays_out, float* __restrict__ target_out,
                                  int* __restrict__ pix_out) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int i, j;
    if (pix_in) {
        i = pix_in[idx] / W; j = pix_in[idx] % W;
    } else {
        const uint32_t domain = (uint32_t)crop_h * (uint32_t)crop_w;
        uint32_t x = (uint32_t)idx;
        do { x = feistel(x, half_bits, key); } while (x >= domain);
        i = r0 + (int)(x / (uint32_t)crop_w); j = c0 + (int)(x % (uint32_t)crop_w);
    }
    // main.py:36-42 (same roundings as compute_rays: the batch equals a gather from the full grid)
    float R[3][3], t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        R[a][0] = __ldg(pose + 4 * a); R[a][1] = __ldg(pose + 4 * a + 1); R[a][2] = __ldg(pose + 4 * a + 2);
        t[a] = __ldg(pose + 4 * a + 3);
    }
    const float dx = __fdiv_rn(__fsub_rn((float)j, half_w), f);
    const float dy = __fdiv_rn(-__fsub_rn((float)i, half_h), f);
    float o[3] = {t[0], t[1], t[2]}, d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        d[a] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[a][0]), __fmul_rn(dy, R[a][1])), -R[a][2]);
    float nn = __fmul_rn(d[0], d[0]);
    nn = __fmaf_rn(d[1], d[1], nn);
    nn = __fmaf_rn(d[2], d[2], nn);
    const float nrm = __fsqrt_rn(nn);
    const float v0 = __fdiv_rn(d[0], nrm), v1 = __fdiv_rn(d[1], nrm), v2 = __fdiv_rn(d[2], nrm);
    if (ndc) {   // data_helpers.py:327-344 with near plane 1 (main.py:68), quirks included
        const float tm = __fdiv_rn(-__fadd_rn(1.f, o[2]), d[2]);
        const float ox = __fadd_rn(o[0], __fmul_rn(tm, o[0]));
        const float oy = __fadd_rn(o[1], __fmul_rn(tm, o[1]));
        const float oz = __fadd_rn(o[2], __fmul_rn(tm, o[2]));
        const float O0 = __fdiv_rn(__fmul_rn(cw, ox), oz);
        const float O1 = __fdiv_rn(__fmul_rn(ch, oy), oz);
        const float O2 = __fadd_rn(1.f, __fdiv_rn(2.f, oz));
        const float D0 = __fmul_rn(cw, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(O0, O2)));
        const float D1 = __fmul_rn(ch, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(O1, O2)));
        const float D2 = __fdiv_rn(-2.f, O2);
        o[0] = O0; o[1] = O1; o[2] = O2; d[0] = D0; d[1] = D1; d[2] = D2;
    }
    float* r = rays_out + NERF_RAY_STRIDE * idx;
    r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
    r[3] = d[0]; r[4] = d[1]; r[5] = d[2];
    r[6] = near; r[7] = far;
    r[8] = v0; r[9] = v1; r[10] = v2;
    if (target_out && image) {
        const float* px = image + ((size_t)i * W + j) * 3;
        target_out[3 * idx] = __ldg(px); target_out[3 * idx + 1] = __ldg(px + 1); target_out[3 * idx + 2] = __ldg(px + 2);
    }
    if (pix_out) pix_out[idx] = i * W + j;
}

int fill_scalars(AdamScalars& s, float lr, float beta1, float beta2, float eps, long step, float grad_scale) {
    if (step < 1) return nerf::arg_error("adam: step must be >= 1");
    s.lr = lr; s.beta1 = beta1; s.beta2 = beta2; s.eps = eps; s.grad_scale = grad_scale;
    s.bias_correction1 = (float)(1.0 - pow((double)beta1, (double)step));
    s.bias_correction2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    return 0;
}

}  // namespace

extern "C" int nerf_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const long* sizes, float lr, float beta1, float beta2,
                              float eps, long step, float grad_scale, void* stream) {
    nerf::DeviceGuard device_guard((n_tensors > 0 && params ? params[0] : nullptr));
    if (n_tensors < 0 || (n_tensors > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !sizes)))
        return nerf::arg_error("nerf_adam_step");
    AdamScalars s;
    int rc = fill_scalars(s, lr, beta1, beta2, eps, step, grad_scale);
    if (rc) return rc;
    for (int base = 0; base < n_tensors; base += kMaxTensors) {
        AdamTensors T;
        const int cnt = n_tensors - base < kMaxTensors ? n_tensors - base : kMaxTensors;
        for (int i = 0; i < cnt; ++i) {
            T.p[i] = params[base + i]; T.g[i] = grads[base + i]; T.m[i] = exp_avg[base + i]; T.v[i] = exp_avg_sq[base + i];
            if (sizes[base + i] < 0 || sizes[base + i] > 0x7fffffffL) return nerf::arg_error("nerf_adam_step: tensor size");
            T.n[i] = (int)sizes[base + i];
            if (T.n[i] > 0 && (!T.p[i] || !T.g[i] || !T.m[i] || !T.v[i])) return nerf::arg_error("nerf_adam_step: null tensor");
        }
        adam_tensors_kernel<<<dim3(16, cnt), 256, 0, (cudaStream_t)stream>>>(T, s);
        rc = nerf::check_launch("nerf_adam_step");
        if (rc) return rc;
    }
    return 0;
}

extern "C" int nerf_adam_step_blob(const float* grad_blob, float* const* params, float* const* exp_avg,
                                   float* const* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                                   long step, float grad_scale, void* stream) {
    nerf::DeviceGuard device_guard(grad_blob);
    if (!grad_blob || !params || !exp_avg || !exp_avg_sq) return nerf::arg_error("nerf_adam_step_blob");
    AdamScalars s;
    int rc = fill_scalars(s, lr, beta1, beta2, eps, step, grad_scale);
    if (rc) return rc;
    BlobTensors T;
    for (int i = 0; i < NERF_N_PARAM_TENSORS; ++i) {
        T.p[i] = params[i]; T.m[i] = exp_avg[i]; T.v[i] = exp_avg_sq[i];
        if (!T.p[i] || !T.m[i] || !T.v[i]) return nerf::arg_error("nerf_adam_step_blob: null tensor");
    }
    adam_blob_kernel<<<dim3(16, NERF_N_PARAM_TENSORS), 256, 0, (cudaStream_t)stream>>>(grad_blob, T, s);
    return nerf::check_launch("nerf_adam_step_blob");
}

extern "C" int nerf_adam_step_blob_peers(const float* const* peer_blobs, int world, float* const* params,
                                         float* const* exp_avg, float* const* exp_avg_sq, float lr, float beta1,
                                         float beta2, float eps, long step, float grad_scale, void* stream) {
    nerf::DeviceGuard device_guard((params ? params[0] : nullptr));
    if (!peer_blobs || world < 1 || world > kMaxPeers || !params || !exp_avg || !exp_avg_sq)
        return nerf::arg_error("nerf_adam_step_blob_peers");
    AdamScalars s;
    int rc = fill_scalars(s, lr, beta1, beta2, eps, step, grad_scale);
    if (rc) return rc;
    PeerBlobs B;
    B.world = world;
    for (int k = 0; k < world; ++k) {
        B.blob[k] = peer_blobs[k];
        if (!B.blob[k]) return nerf::arg_error("nerf_adam_step_blob_peers: null peer pointer");
    }
    BlobTensors T;
    for (int i = 0; i < NERF_N_PARAM_TENSORS; ++i) {
        T.p[i] = params[i]; T.m[i] = exp_avg[i]; T.v[i] = exp_avg_sq[i];
        if (!T.p[i] || !T.m[i] || !T.v[i]) return nerf::arg_error("nerf_adam_step_blob_peers: null tensor");
    }
    adam_blob_peers_kernel<<<dim3(16, NERF_N_PARAM_TENSORS), 256, 0, (cudaStream_t)stream>>>(B, T, s);
    return nerf::check_launch("nerf_adam_step_blob_peers");
}

extern "C" int nerf_train_rays(int H, int W, float focal, float cw, float ch, const float* pose, const int* pix,
                               unsigned long long seed, int crop_r0, int crop_c0, int crop_h, int crop_w, long n,
                               int ndc, float near, float far, const float* image, float* rays_out,
                               float* target_out, int* pix_out, void* stream) {
    nerf::DeviceGuard device_guard(rays_out);
    if (H <= 0 || W <= 0 || !pose || n < 0 || (n > 0 && !rays_out)) return nerf::arg_error("nerf_train_rays");
    if (n == 0) return 0;
    int half_bits = 1;
    if (!pix) {
        if (crop_r0 < 0 || crop_c0 < 0 || crop_h <= 0 || crop_w <= 0 || crop_r0 + crop_h > H || crop_c0 + crop_w > W)
            return nerf::arg_error("nerf_train_rays: crop window");
        const long domain = (long)crop_h * crop_w;
        if (n > domain) return nerf::arg_error("nerf_train_rays: more rays than pixels (sampling is without replacement)");
        while ((1L << (2 * half_bits)) < domain) ++half_bits;
        if (half_bits > 15) return nerf::arg_error("nerf_train_rays: image too large");
    }
    train_rays_kernel<<<nerf::blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        H, W, (float)(H * .5), (float)(W * .5), focal, cw, ch, pose, pix, (uint64_t)seed, crop_r0, crop_c0, crop_h, crop_w,
        half_bits, n, ndc, near, far, image, rays_out, target_out, pix_out);
    return nerf::check_launch("nerf_train_rays");
}
