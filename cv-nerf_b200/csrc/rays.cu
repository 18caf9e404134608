// K1: ray generation, NDC warp, ray packing and coarse depth sampling.
//
// Everything here is fp32 with the reference's operation order and one rounding per
// operation (explicit __f*_rn intrinsics are never contracted into FMAs), so results are
// bit-identical to the reference's torch-CPU path:
//   compute_rays   /root/reference/main.py:19-46
//   get_ndc        /root/reference/data_helpers.py:327-344
//   render() front /root/reference/main.py:55-76
//   coarse z       /root/reference/main.py:221-234
// The kernels are HBM-write bound (44 B/ray); one thread per ray, rays_out written as 11
// consecutive floats per thread (consecutive threads -> consecutive 44 B records).
#include "common.cuh"
#include "rng.cuh"

namespace {

struct Pose {
    float r[3][3];
    float t[3];
};

__device__ __forceinline__ Pose load_pose(const float* __restrict__ p) {
    Pose q;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q.r[a][0] = __ldg(p + 4 * a + 0);
        q.r[a][1] = __ldg(p + 4 * a + 1);
        q.r[a][2] = __ldg(p + 4 * a + 2);
        q.t[a] = __ldg(p + 4 * a + 3);
    }
    return q;
}

// main.py:36-42: d = ((j - W/2)/f, -(i - H/2)/f, -1), world = ((d0*R0 + d1*R1) + d2*R2)
__device__ __forceinline__ void pixel_dir(int i, int j, float half_h, float half_w, float f,
                                          const Pose& P, float d[3]) {
    float dx = __fdiv_rn(__fsub_rn((float)j, half_w), f);
    float dy = __fdiv_rn(-__fsub_rn((float)i, half_h), f);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float s = __fadd_rn(__fmul_rn(dx, P.r[a][0]), __fmul_rn(dy, P.r[a][1]));
        d[a] = __fadd_rn(s, -P.r[a][2]);  // (-1) * R is exact
    }
}

// data_helpers.py:327-344 with the reference's two quirks (origin shifted by t*o; directions use
// the already-warped origin) and left-to-right evaluation.
__device__ __forceinline__ void ndc_warp(float cw, float ch, float near_plane, float o[3],
                                         float d[3]) {
    float tm = __fdiv_rn(-__fadd_rn(near_plane, o[2]), d[2]);
    float ox = __fadd_rn(o[0], __fmul_rn(tm, o[0]));
    float oy = __fadd_rn(o[1], __fmul_rn(tm, o[1]));
    float oz = __fadd_rn(o[2], __fmul_rn(tm, o[2]));
    float two_near = __fmul_rn(2.f, near_plane);
    float O0 = __fdiv_rn(__fmul_rn(cw, ox), oz);
    float O1 = __fdiv_rn(__fmul_rn(ch, oy), oz);
    float O2 = __fadd_rn(1.f, __fdiv_rn(two_near, oz));
    float D0 = __fmul_rn(cw, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(O0, O2)));
    float D1 = __fmul_rn(ch, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(O1, O2)));
    float D2 = __fdiv_rn(-two_near, O2);
    o[0] = O0; o[1] = O1; o[2] = O2;
    d[0] = D0; d[1] = D1; d[2] = D2;
}

__global__ void compute_rays_kernel(int W, float half_h, float half_w, float f,
                                    const float* __restrict__ pose, int row0, long n,
                                    float* __restrict__ origins, float* __restrict__ dirs) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    Pose P = load_pose(pose);
    int i = row0 + (int)(idx / W), j = (int)(idx % W);
    float d[3];
    pixel_dir(i, j, half_h, half_w, f, P, d);
    dirs[3 * idx + 0] = d[0]; dirs[3 * idx + 1] = d[1]; dirs[3 * idx + 2] = d[2];
    if (origins) {
        origins[3 * idx + 0] = P.t[0]; origins[3 * idx + 1] = P.t[1]; origins[3 * idx + 2] = P.t[2];
    }
}

__global__ void get_ndc_kernel(float cw, float ch, float near_plane, const float* __restrict__ o_in,
                               const float* __restrict__ d_in, long n, float* __restrict__ o_out,
                               float* __restrict__ d_out) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float o[3] = {o_in[3 * idx], o_in[3 * idx + 1], o_in[3 * idx + 2]};
    float d[3] = {d_in[3 * idx], d_in[3 * idx + 1], d_in[3 * idx + 2]};
    ndc_warp(cw, ch, near_plane, o, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) { o_out[3 * idx + a] = o[a]; d_out[3 * idx + a] = d[a]; }
}

__global__ void pack_rays_kernel(int W, float half_h, float half_w, float f, float cw, float ch,
                                 const float* __restrict__ pose, int row0,
                                 const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                 long n, int ndc, float near, float far, float* __restrict__ out) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float o[3], d[3];
    if (pose) {
        Pose P = load_pose(pose);
        pixel_dir(row0 + (int)(idx / W), (int)(idx % W), half_h, half_w, f, P, d);
        o[0] = P.t[0]; o[1] = P.t[1]; o[2] = P.t[2];
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) { o[a] = rays_o[3 * idx + a]; d[a] = rays_d[3 * idx + a]; }
    }
    // main.py:61-63: torch.norm over 3 elements accumulates with fused multiply-adds on CPU
    float nn = __fmul_rn(d[0], d[0]);
    nn = __fmaf_rn(d[1], d[1], nn);
    nn = __fmaf_rn(d[2], d[2], nn);
    float nrm = __fsqrt_rn(nn);
    float v0 = __fdiv_rn(d[0], nrm), v1 = __fdiv_rn(d[1], nrm), v2 = __fdiv_rn(d[2], nrm);
    if (ndc) ndc_warp(cw, ch, 1.f, o, d);  // main.py:68 passes near = 1.
    float* r = out + NERF_RAY_STRIDE * idx;
    r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
    r[3] = d[0]; r[4] = d[1]; r[5] = d[2];
    r[6] = near; r[7] = far;
    r[8] = v0; r[9] = v1; r[10] = v2;
}

// torch.linspace(0, 1, S)[i] on CPU: step*i below the midpoint, fma(-step, S-1-i, 1) above it.
__device__ __forceinline__ float unit_linspace(int i, int S, float step) {
    return (i < S / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(S - 1 - i), 1.f);
}

__device__ __forceinline__ float coarse_z(int i, int S, float step, float near, float far) {
    float s = unit_linspace(i, S, step);
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, s)), __fmul_rn(far, s));
}

// One thread per (ray, group of four consecutive samples): 32-bit index arithmetic and, when S % 4 == 0,
// one 16-byte load of the jitter and one 16-byte store of the depths (the first version ran one thread per
// sample with a 64-bit divide: 137 us per 800x800 frame at 75 % issue, ncu profiles/r02_small_kernels_ncu.txt).
__global__ void sample_coarse_kernel(const float* __restrict__ rays, long n, int S,
                                     const float* __restrict__ t_rand, float* __restrict__ z_out) {
    const int groups = (S + 3) / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * groups) return;
    const long ray = idx / groups;
    const int g = (int)(idx - ray * groups);
    const float near = __ldg(rays + NERF_RAY_STRIDE * ray + 6), far = __ldg(rays + NERF_RAY_STRIDE * ray + 7);
    const float step = __fdiv_rn(1.f, (float)(S - 1));
    const long base = ray * S + g * 4;
    const bool vec = (S & 3) == 0;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (t_rand) {
        if (vec) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(t_rand + base));
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (g * 4 + j < S) t[j] = __ldg(t_rand + base + j);
        }
    }
    float z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = g * 4 + j;
        float v = coarse_z(min(i, S - 1), S, step, near, far);
        if (t_rand && i < S) {  // main.py:227-234
            float lo = v, hi = v;
            if (i > 0) lo = __fmul_rn(.5f, __fadd_rn(v, coarse_z(i - 1, S, step, near, far)));
            if (i < S - 1) hi = __fmul_rn(.5f, __fadd_rn(coarse_z(i + 1, S, step, near, far), v));
            v = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t[j]));
        }
        z[j] = v;
    }
    if (vec) {
        *reinterpret_cast<float4*>(z_out + base) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (g * 4 + j < S) z_out[base + j] = z[j];
    }
}

// The same with the jitter drawn in place: one thread per (ray, group of four samples).
__global__ void sample_coarse_rng_kernel(const float* __restrict__ rays, long n, int S, unsigned long long seed,
                                         long ray0, float* __restrict__ z_out) {
    const int groups = (S + 3) / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * groups) return;
    const long ray = idx / groups;
    const int g = (int)(idx % groups);
    const float near = __ldg(rays + NERF_RAY_STRIDE * ray + 6), far = __ldg(rays + NERF_RAY_STRIDE * ray + 7);
    const float step = __fdiv_rn(1.f, (float)(S - 1));
    const float4 t = nerf::uniform4(nerf::draw_group(seed, NERF_RNG_STREAM_T_RAND, ray0 + ray, g));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = g * 4 + j;
        if (i >= S) break;
        const float z = coarse_z(i, S, step, near, far);
        float lo = z, hi = z;
        if (i > 0) lo = __fmul_rn(.5f, __fadd_rn(z, coarse_z(i - 1, S, step, near, far)));
        if (i < S - 1) hi = __fmul_rn(.5f, __fadd_rn(coarse_z(i + 1, S, step, near, far), z));
        z_out[ray * S + i] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), nerf::pick4(t, j)));
    }
}

// to_byte / cont_to_byte8_im (model.py:134, utils.py:57): (255 * clip(x, 0, 1)).astype(uint8);
// numpy multiplies in fp32 and astype truncates toward zero.  4 values per thread.
__global__ void to_byte_kernel(const float* __restrict__ x, long n, uint8_t* __restrict__ out) {
    long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n && ((uintptr_t)(x + i) & 15) == 0 && ((uintptr_t)(out + i) & 3) == 0) {
        const float4 v = *reinterpret_cast<const float4*>(x + i);
        uchar4 o;
        o.x = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(v.x, 0.f), 1.f));
        o.y = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(v.y, 0.f), 1.f));
        o.z = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(v.z, 0.f), 1.f));
        o.w = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(v.w, 0.f), 1.f));
        *reinterpret_cast<uchar4*>(out + i) = o;
    } else {
        for (long k = i; k < n && k < i + 4; ++k) out[k] = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(x[k], 0.f), 1.f));
    }
}

}  // namespace

extern "C" int nerf_to_byte(const float* x, long n, unsigned char* out, void* stream) {
    nerf::DeviceGuard device_guard(out);
    if (n < 0 || (n > 0 && (!x || !out))) return nerf::arg_error("nerf_to_byte");
    if (n == 0) return 0;
    to_byte_kernel<<<nerf::blocks_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, n, out);
    return nerf::check_launch("nerf_to_byte");
}

extern "C" int nerf_compute_rays(int H, int W, float focal, const float* pose, int row0, int row1,
                                 float* origins_out, float* dirs_out, void* stream) {
    nerf::DeviceGuard device_guard(dirs_out);
    if (H <= 0 || W <= 0 || !pose || !dirs_out || row0 < 0 || row1 > H || row0 > row1)
        return nerf::arg_error("nerf_compute_rays");
    long n = (long)(row1 - row0) * W;
    if (n == 0) return 0;
    compute_rays_kernel<<<nerf::blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        W, (float)(H * .5), (float)(W * .5), focal, pose, row0, n, origins_out, dirs_out);
    return nerf::check_launch("nerf_compute_rays");
}

extern "C" int nerf_get_ndc(float cw, float ch, float near_plane, const float* o, const float* d,
                            long n, float* o_out, float* d_out, void* stream) {
    nerf::DeviceGuard device_guard(o_out);
    if (n < 0 || (n > 0 && (!o || !d || !o_out || !d_out))) return nerf::arg_error("nerf_get_ndc");
    if (n == 0) return 0;
    get_ndc_kernel<<<nerf::blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(cw, ch, near_plane, o, d,
                                                                            n, o_out, d_out);
    return nerf::check_launch("nerf_get_ndc");
}

extern "C" int nerf_pack_rays(int H, int W, float focal, float cw, float ch, const float* pose,
                              int row0, int row1, const float* rays_o, const float* rays_d, long n,
                              int ndc, float near, float far, float* rays_out, void* stream) {
    nerf::DeviceGuard device_guard(rays_out);
    if (pose) {
        if (H <= 0 || W <= 0 || row0 < 0 || row1 > H || row0 > row1) return nerf::arg_error("nerf_pack_rays rows");
        n = (long)(row1 - row0) * W;
    } else if (!rays_o || !rays_d) {
        if (n != 0) return nerf::arg_error("nerf_pack_rays: need pose or rays");
    }
    if (n < 0 || (n > 0 && !rays_out)) return nerf::arg_error("nerf_pack_rays");
    if (n == 0) return 0;
    pack_rays_kernel<<<nerf::blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        W, (float)(H * .5), (float)(W * .5), focal, cw, ch, pose, row0, rays_o, rays_d, n, ndc, near, far,
        rays_out);
    return nerf::check_launch("nerf_pack_rays");
}

extern "C" int nerf_sample_coarse_rng(const float* rays, long n, int S, unsigned long long seed, long ray0,
                                      float* z_out, void* stream) {
    nerf::DeviceGuard device_guard(z_out);
    if (n < 0 || S < 2 || (n > 0 && (!rays || !z_out))) return nerf::arg_error("nerf_sample_coarse_rng");
    if (n == 0) return 0;
    sample_coarse_rng_kernel<<<nerf::blocks_for(n * ((S + 3) / 4), 256), 256, 0, (cudaStream_t)stream>>>(rays, n, S, seed,
                                                                                                     ray0, z_out);
    return nerf::check_launch("nerf_sample_coarse_rng");
}

extern "C" int nerf_sample_coarse(const float* rays, long n, int S, const float* t_rand,
                                  float* z_out, void* stream) {
    nerf::DeviceGuard device_guard(z_out);
    if (n < 0 || S < 2 || (n > 0 && (!rays || !z_out))) return nerf::arg_error("nerf_sample_coarse");
    if (n == 0) return 0;
    sample_coarse_kernel<<<nerf::blocks_for(n * ((S + 3) / 4), 256), 256, 0, (cudaStream_t)stream>>>(rays, n, S, t_rand,
                                                                                                 z_out);
    return nerf::check_launch("nerf_sample_coarse");
}
