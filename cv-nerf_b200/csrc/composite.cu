// K3/K4: alpha compositing (forward + backward), inverse-CDF resampling and sort-merge.
//
// Reference sites: process_volume_info /root/reference/main.py:170-204,
//                  inv_transform_sampling /root/reference/utils.py:4-53,
//                  midpoints + sort-merge /root/reference/main.py:248-251.
//
// One warp owns one ray.  A lane holds C = ceil(S/32) *consecutive* samples so the exclusive
// cumprod (transmittance) is a local product, one 5-step shuffle scan over the lane products and
// a local sweep.  The kernels are HBM-bound (16 B raw + 4 B z per sample in, 12 B per ray out)
// and together account for <1% of a render; blocks of 8 warps, grid = ceil(n/8).
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr float FAR_DELTA = 1e10f;      // main.py:175
constexpr float TRANSMIT_EPS = 1e-10f;  // main.py:194
constexpr float PDF_EPS = 1e-5f;        // utils.py:12
constexpr int MAX_SORT = 256;

__device__ __forceinline__ float dir_norm(const float* __restrict__ d) {
    // torch.norm over 3 elements on CPU: x*x, then two fused multiply-adds, then sqrt
    float x = __ldg(d), y = __ldg(d + 1), z = __ldg(d + 2);
    float nn = __fmul_rn(x, x);
    nn = __fmaf_rn(y, y, nn);
    nn = __fmaf_rn(z, z, nn);
    return __fsqrt_rn(nn);
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_excl_scan_mul(float v, int lane) {
    float incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl *= o;
    }
    float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 1.f : ex;
}

__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
    float incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float o = __shfl_down_sync(0xffffffffu, incl, off);
        if (lane + off < 32) incl += o;
    }
    float ex = __shfl_down_sync(0xffffffffu, incl, 1);
    return lane == 31 ? 0.f : ex;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Per-sample forward terms shared by forward and backward.
struct SampleTerms {
    float alpha, t, delta, pre;  // t = 1 - alpha + eps; pre = sigma + noise
};

__device__ __forceinline__ SampleTerms sample_terms(float sigma, float noise, float z0, float z1,
                                                    bool last, float nrm) {
    SampleTerms s;
    float dz = last ? FAR_DELTA : (z1 - z0);
    s.delta = dz * nrm;
    s.pre = sigma + noise;
    s.alpha = 1.f - expf(s.delta * -fmaxf(s.pre, 0.f));
    s.t = 1.f - s.alpha + TRANSMIT_EPS;
    return s;
}

template <int C>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
composite_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ dirs, int dir_stride, const float* __restrict__ noise,
                     long n, int S, int white_bkg, float* __restrict__ rgb_out,
                     float* __restrict__ weights_out) {
    const int lane = threadIdx.x & 31;
    const long ray = (long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (ray >= n) return;
    const float nrm = dir_norm(dirs + ray * dir_stride);
    const long base = ray * S;
    float alpha[C], rgb[C][3];
    float prod = 1.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = lane * C + c;
        alpha[c] = 0.f;
        rgb[c][0] = rgb[c][1] = rgb[c][2] = 0.f;
        if (i < S) {
            float4 r = __ldg(raw + base + i);
            float z0 = __ldg(z + base + i);
            float z1 = (i + 1 < S) ? __ldg(z + base + i + 1) : 0.f;
            float nz = noise ? __ldg(noise + base + i) : 0.f;
            SampleTerms s = sample_terms(r.w, nz, z0, z1, i + 1 == S, nrm);
            alpha[c] = s.alpha;
            prod *= s.t;
            rgb[c][0] = sigmoidf(r.x); rgb[c][1] = sigmoidf(r.y); rgb[c][2] = sigmoidf(r.z);
        }
    }
    float T = warp_excl_scan_mul(prod, lane);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accw = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = lane * C + c;
        if (i < S) {
            float w = alpha[c] * T;
            T *= (1.f - alpha[c] + TRANSMIT_EPS);
            acc0 += w * rgb[c][0]; acc1 += w * rgb[c][1]; acc2 += w * rgb[c][2];
            accw += w;
            if (weights_out) weights_out[base + i] = w;
        }
    }
    acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2); accw = warp_sum(accw);
    if (lane == 0) {
        float bg = white_bkg ? (1.f - accw) : 0.f;
        rgb_out[3 * ray + 0] = acc0 + bg;
        rgb_out[3 * ray + 1] = acc1 + bg;
        rgb_out[3 * ray + 2] = acc2 + bg;
    }
}

// Closed-form backward:
//   G_i        = g . rgb_i - [bkg] sum(g) + gw_i
//   dL/dalpha_i = T_i G_i - (sum_{j>i} w_j G_j) / t_i
//   dL/dsigma_i = dL/dalpha_i * delta_i * (1 - alpha_i) * [sigma_i + noise_i > 0]
//   dL/drgbraw_i = w_i * g * rgb_i (1 - rgb_i)
template <int C>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
composite_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ dirs, int dir_stride, const float* __restrict__ noise,
                     long n, int S, int white_bkg, const float* __restrict__ grad_rgb,
                     const float* __restrict__ grad_w, float4* __restrict__ grad_raw) {
    const int lane = threadIdx.x & 31;
    const long ray = (long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (ray >= n) return;
    const float nrm = dir_norm(dirs + ray * dir_stride);
    const long base = ray * S;
    const float g0 = __ldg(grad_rgb + 3 * ray), g1 = __ldg(grad_rgb + 3 * ray + 1),
                g2 = __ldg(grad_rgb + 3 * ray + 2);
    const float gbg = white_bkg ? (g0 + g1 + g2) : 0.f;
    SampleTerms st[C];
    float rgb[C][3], G[C];
    float prod = 1.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = lane * C + c;
        st[c].alpha = 0.f; st[c].t = 1.f; st[c].delta = 0.f; st[c].pre = 0.f;
        rgb[c][0] = rgb[c][1] = rgb[c][2] = 0.f;
        G[c] = 0.f;
        if (i < S) {
            float4 r = __ldg(raw + base + i);
            float z0 = __ldg(z + base + i);
            float z1 = (i + 1 < S) ? __ldg(z + base + i + 1) : 0.f;
            float nz = noise ? __ldg(noise + base + i) : 0.f;
            st[c] = sample_terms(r.w, nz, z0, z1, i + 1 == S, nrm);
            prod *= st[c].t;
            rgb[c][0] = sigmoidf(r.x); rgb[c][1] = sigmoidf(r.y); rgb[c][2] = sigmoidf(r.z);
            G[c] = g0 * rgb[c][0] + g1 * rgb[c][1] + g2 * rgb[c][2] - gbg +
                   (grad_w ? __ldg(grad_w + base + i) : 0.f);
        }
    }
    float T = warp_excl_scan_mul(prod, lane);
    float w[C], Tc[C];
    float lane_sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        Tc[c] = T;
        w[c] = st[c].alpha * T;
        T *= st[c].t;
        lane_sum += w[c] * G[c];
    }
    float suffix = warp_excl_suffix_sum(lane_sum, lane);  // sum over later lanes
#pragma unroll
    for (int c = C - 1; c >= 0; --c) {
        int i = lane * C + c;
        if (i < S) {
            float dalpha = Tc[c] * G[c] - suffix / st[c].t;
            float dsigma = (st[c].pre > 0.f) ? dalpha * st[c].delta * (1.f - st[c].alpha) : 0.f;
            float4 o;
            o.x = w[c] * g0 * rgb[c][0] * (1.f - rgb[c][0]);
            o.y = w[c] * g1 * rgb[c][1] * (1.f - rgb[c][1]);
            o.z = w[c] * g2 * rgb[c][2] * (1.f - rgb[c][2]);
            o.w = dsigma;
            grad_raw[base + i] = o;
        }
        suffix += w[c] * G[c];
    }
}

// Optional extra maps from the compositing weights (the reference computes only sum(w) internally,
// main.py:199, and returns none of them; these are nerf-pytorch raw2outputs' definitions, which the
// reference's README lists as the model it follows): depth = sum w z, acc = sum w,
// disp = 1 / max(1e-10, depth / acc).  One warp per ray.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
composite_maps_kernel(const float* __restrict__ w, const float* __restrict__ z, long n, int S,
                      float* __restrict__ maps) {
    const int lane = threadIdx.x & 31;
    const long ray = (long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (ray >= n) return;
    float depth = 0.f, acc = 0.f;
    for (int i = lane; i < S; i += 32) {
        const float wi = __ldg(w + ray * S + i);
        depth = fmaf(wi, __ldg(z + ray * S + i), depth);
        acc += wi;
    }
    depth = warp_sum(depth); acc = warp_sum(acc);
    if (lane == 0) {
        maps[3 * ray + 0] = depth;
        maps[3 * ray + 1] = acc;
        maps[3 * ray + 2] = 1.f / fmaxf(1e-10f, depth / acc);
    }
}

// ---------------------------------------------------------------------------------- resampling

// Monotone map float -> uint32 (NaNs sort last, like torch.sort).
__device__ __forceinline__ uint32_t order_key(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Build cdf[0..B-1] (cdf[0] = 0) in shared memory from B-1 weights; all 32 lanes participate.
__device__ __forceinline__ void build_cdf(const float* __restrict__ w, int nw, float* cdf, int lane) {
    // pdf = (w + eps) / sum(w + eps); cdf = cumsum(pdf).  Lane-contiguous chunks + shuffle scan.
    const int per = (nw + 31) / 32;
    float local = 0.f;
    for (int c = 0; c < per; ++c) {
        int i = lane * per + c;
        if (i < nw) local += __ldg(w + i) + PDF_EPS;
    }
    float total = warp_sum(local);
    // exclusive prefix of lane sums (in pdf units)
    float incl = local / total;
    float lane_pdf = incl;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    float run = incl - lane_pdf;
    if (lane == 0) cdf[0] = 0.f;
    for (int c = 0; c < per; ++c) {
        int i = lane * per + c;
        if (i < nw) {
            run += (__ldg(w + i) + PDF_EPS) / total;
            cdf[i + 1] = run;
        }
    }
}

// utils.py:35-53 for one u.  cdf/bins have B entries.
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int B, float u) {
    int lo_i = 0, hi_i = B;  // first index with cdf[idx] > u  (searchsorted right=True)
    while (lo_i < hi_i) {
        int mid = (lo_i + hi_i) >> 1;
        if (cdf[mid] <= u) lo_i = mid + 1; else hi_i = mid;
    }
    int idx = lo_i;
    int lower = max(idx - 1, 0), upper = min(idx, B - 1);
    float c0 = cdf[lower], c1 = cdf[upper];
    float b0 = bins[lower], b1 = bins[upper];
    float span = c1 - c0;
    if (span < PDF_EPS) span = 1.f;
    return (b1 - b0) * ((u - c0) / span) + b0;
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                  const float* __restrict__ u, long n, int B, int m, float* __restrict__ out) {
    __shared__ float s_cdf[WARPS_PER_BLOCK][MAX_SORT];
    __shared__ float s_bins[WARPS_PER_BLOCK][MAX_SORT];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long ray = (long)blockIdx.x * WARPS_PER_BLOCK + wib;
    if (ray >= n) return;
    for (int i = lane; i < B; i += 32) s_bins[wib][i] = __ldg(bins + ray * B + i);
    build_cdf(weights + ray * (B - 1), B - 1, s_cdf[wib], lane);
    __syncwarp();
    for (int k = lane; k < m; k += 32)
        out[ray * m + k] = invert_cdf(s_cdf[wib], s_bins[wib], B, __ldg(u + ray * m + k));
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
resample_merge_kernel(const float* __restrict__ z_c, const float* __restrict__ w_c,
                      const float* __restrict__ u, long n, int S, int m, float* __restrict__ z_f) {
    __shared__ __align__(16) float s_cdf[WARPS_PER_BLOCK][MAX_SORT];
    __shared__ float s_bins[WARPS_PER_BLOCK][MAX_SORT];
    __shared__ __align__(16) uint32_t s_key[WARPS_PER_BLOCK][MAX_SORT];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long ray = (long)blockIdx.x * WARPS_PER_BLOCK + wib;
    if (ray >= n) return;
    const int B = S - 1, total = S + m;
    const float* zr = z_c + ray * S;
    for (int i = lane; i < S; i += 32) {
        float zi = __ldg(zr + i);
        s_key[wib][i] = order_key(zi);
        if (i < B) s_bins[wib][i] = .5f * (__ldg(zr + i + 1) + zi);  // main.py:248
    }
    build_cdf(w_c + ray * S + 1, S - 2, s_cdf[wib], lane);           // weights[..., 1:-1]
    __syncwarp();
    for (int k = lane; k < m; k += 32)
        s_key[wib][S + k] = order_key(invert_cdf(s_cdf[wib], s_bins[wib], B, __ldg(u + ray * m + k)));
    __syncwarp();
    // Values-only sort of [coarse depths | new samples] (main.py:251).  The coarse depths are already
    // ascending in every configuration of the reference (linspace, stratified jitter keeps the
    // order): then only the m new samples need ranking among themselves (m^2/32 compares per lane,
    // four keys per shared-memory load) and the two sorted lists are merged by binary search.
    // Ties: coarse first (strict / non-strict counts), which is a permutation and gives the same
    // VALUES as any other tie order.  Unsorted coarse depths take the general rank sort.
    bool sorted = true;
    for (int i = lane; i + 1 < S; i += 32) sorted &= s_key[wib][i] <= s_key[wib][i + 1];
    sorted = __all_sync(0xffffffffu, sorted);
    float* out = z_f + ray * total;
    auto decode = [](uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); };
    if (sorted && (m & 3) == 0 && (S & 3) == 0) {
        uint32_t* samp = s_key[wib] + S;                     // m keys, 16-byte aligned (S % 4 == 0)
        uint32_t* ranked = reinterpret_cast<uint32_t*>(s_cdf[wib]);   // cdf no longer needed
        for (int i = lane; i < m; i += 32) {
            const uint32_t ki = samp[i];
            int rank = 0;
            for (int j = 0; j < m; j += 4) {
                const uint4 q = *reinterpret_cast<const uint4*>(samp + j);
                rank += (q.x < ki) || (q.x == ki && j + 0 < i);
                rank += (q.y < ki) || (q.y == ki && j + 1 < i);
                rank += (q.z < ki) || (q.z == ki && j + 2 < i);
                rank += (q.w < ki) || (q.w == ki && j + 3 < i);
            }
            ranked[rank] = ki;
        }
        __syncwarp();
        for (int i = lane; i < total; i += 32) {
            if (i < S) {                                     // coarse i: + #samples strictly below
                const uint32_t k = s_key[wib][i];
                int lo = 0, hi = m;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (ranked[mid] < k) lo = mid + 1; else hi = mid; }
                out[i + lo] = decode(k);
            } else {                                         // sample of rank r: + #coarse at or below
                const int r = i - S;
                const uint32_t k = ranked[r];
                int lo = 0, hi = S;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (s_key[wib][mid] <= k) lo = mid + 1; else hi = mid; }
                out[r + lo] = decode(k);
            }
        }
        return;
    }
    // general rank sort: rank = #{j : key_j < key_i or (key_j == key_i and j < i)}
    for (int i = lane; i < total; i += 32) {
        uint32_t ki = s_key[wib][i];
        int rank = 0;
        for (int j = 0; j < total; ++j) {
            uint32_t kj = s_key[wib][j];
            rank += (kj < ki) || (kj == ki && j < i);
        }
        out[rank] = decode(ki);
    }
}

template <typename F>
int dispatch_c(int S, F&& f) {
    int C = (S + 31) / 32;
    switch (C) {
        case 1: f(std::integral_constant<int, 1>()); return 0;
        case 2: f(std::integral_constant<int, 2>()); return 0;
        case 3: f(std::integral_constant<int, 3>()); return 0;
        case 4: f(std::integral_constant<int, 4>()); return 0;
        case 5: f(std::integral_constant<int, 5>()); return 0;
        case 6: f(std::integral_constant<int, 6>()); return 0;
        case 7: f(std::integral_constant<int, 7>()); return 0;
        case 8: f(std::integral_constant<int, 8>()); return 0;
        default: return -1;
    }
}

}  // namespace

extern "C" int nerf_composite_fwd(const float* raw, const float* z, const float* dirs,
                                  int dir_stride, const float* noise, long n, int S, int white_bkg,
                                  float* rgb_out, float* weights_out, void* stream) {
    if (n < 0 || S < 1 || (n > 0 && (!raw || !z || !dirs || !rgb_out))) return nerf::arg_error("nerf_composite_fwd");
    if (n == 0) return 0;
    unsigned grid = nerf::blocks_for(n, WARPS_PER_BLOCK);
    int rc = dispatch_c(S, [&](auto c) {
        composite_fwd_kernel<decltype(c)::value><<<grid, WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
            (const float4*)raw, z, dirs, dir_stride, noise, n, S, white_bkg, rgb_out, weights_out);
    });
    if (rc) { nerf::set_last_error("nerf_composite_fwd: S=%d > 256 unsupported", S); return NERF_ERR_UNSUPPORTED; }
    return nerf::check_launch("nerf_composite_fwd");
}

extern "C" int nerf_composite_bwd(const float* raw, const float* z, const float* dirs,
                                  int dir_stride, const float* noise, long n, int S, int white_bkg,
                                  const float* grad_rgb, const float* grad_weights, float* grad_raw,
                                  void* stream) {
    if (n < 0 || S < 1 || (n > 0 && (!raw || !z || !dirs || !grad_rgb || !grad_raw)))
        return nerf::arg_error("nerf_composite_bwd");
    if (n == 0) return 0;
    unsigned grid = nerf::blocks_for(n, WARPS_PER_BLOCK);
    int rc = dispatch_c(S, [&](auto c) {
        composite_bwd_kernel<decltype(c)::value><<<grid, WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
            (const float4*)raw, z, dirs, dir_stride, noise, n, S, white_bkg, grad_rgb, grad_weights,
            (float4*)grad_raw);
    });
    if (rc) { nerf::set_last_error("nerf_composite_bwd: S=%d > 256 unsupported", S); return NERF_ERR_UNSUPPORTED; }
    return nerf::check_launch("nerf_composite_bwd");
}

extern "C" int nerf_sample_pdf(const float* bins, const float* weights, const float* u, long n,
                               int B, int m, float* samples_out, void* stream) {
    if (n < 0 || B < 2 || m < 0 || (n > 0 && m > 0 && (!bins || !weights || !u || !samples_out)))
        return nerf::arg_error("nerf_sample_pdf");
    if (B > MAX_SORT) { nerf::set_last_error("nerf_sample_pdf: B=%d > 256 unsupported", B); return NERF_ERR_UNSUPPORTED; }
    if (n == 0 || m == 0) return 0;
    sample_pdf_kernel<<<nerf::blocks_for(n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        bins, weights, u, n, B, m, samples_out);
    return nerf::check_launch("nerf_sample_pdf");
}

extern "C" int nerf_resample_merge(const float* z_c, const float* w_c, const float* u, long n, int S,
                                   int m, float* z_f, void* stream) {
    if (n < 0 || S < 3 || m < 0 || (n > 0 && (!z_c || !w_c || !z_f || (m > 0 && !u))))
        return nerf::arg_error("nerf_resample_merge");
    if (S + m > MAX_SORT) { nerf::set_last_error("nerf_resample_merge: S+m=%d > 256 unsupported", S + m); return NERF_ERR_UNSUPPORTED; }
    if (n == 0) return 0;
    resample_merge_kernel<<<nerf::blocks_for(n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        z_c, w_c, u, n, S, m, z_f);
    return nerf::check_launch("nerf_resample_merge");
}

extern "C" int nerf_composite_maps(const float* weights, const float* z, long n, int S, float* maps_out,
                                   void* stream) {
    if (n < 0 || S < 1 || (n > 0 && (!weights || !z || !maps_out))) return nerf::arg_error("nerf_composite_maps");
    if (n == 0) return 0;
    composite_maps_kernel<<<nerf::blocks_for(n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        weights, z, n, S, maps_out);
    return nerf::check_launch("nerf_composite_maps");
}
