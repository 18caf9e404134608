// Whole-chain entry: render_rays (/root/reference/main.py:207-261) and, with a pose, the front end of
// render (/root/reference/main.py:49-87) for one batch of rays -- every launch of the chain sequenced on
// the caller's stream by ONE call (`nerf_render_fused`).  The host makes no
// decision between the launches and nothing synchronises; all intermediates live in a caller-provided
// scratch buffer; the random draws are made inside the kernels (csrc/rng.cuh).
//
//   rays (given, or generated from the pose for image rows [row0,row1))
//   -> coarse depths (+ stratified jitter)      nerf_sample_coarse[_rng]
//   -> per-ray view terms of both networks       nerf_viewdir_term x2
//   -> coarse field                              nerf_mlp_fwd_host_tail
//   -> compositing (+ density noise)             nerf_composite_fwd[_rng]
//   -> inverse-CDF resampling + sort-merge       nerf_resample_merge_rng
//   -> fine field                                nerf_mlp_fwd_host_tail
//   -> compositing                               nerf_composite_fwd[_rng]
#include "common.cuh"

namespace {

struct Scratch {
    size_t rays, z_c, raw_c, w_c, z_f, raw_f, vt_c, vt_f, total;
};

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

Scratch scratch_layout(long n, int S_c, int n_fine) {
    Scratch s{};
    const size_t N = (size_t)n, Sf = (size_t)S_c + n_fine;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off += align256(floats * 4); return o; };
    s.rays = take(N * NERF_RAY_STRIDE);
    s.z_c = take(N * S_c);
    s.raw_c = take(N * S_c * 4);
    s.w_c = take(N * S_c);
    s.z_f = take(N * Sf);
    s.raw_f = take(N * Sf * 4);
    s.vt_c = take(N * 128);
    s.vt_f = take(N * 128);
    s.total = off;
    return s;
}

}  // namespace

extern "C" size_t nerf_render_scratch_bytes(long n, int S_c, int n_fine) {
    if (n <= 0 || S_c < 1 || n_fine < 0) return 0;
    return scratch_layout(n, S_c, n_fine).total;
}

extern "C" int nerf_render_fused(const void* packed_coarse, const void* host_tail_coarse, const void* packed_fine,
                                 const void* host_tail_fine, int H, int W, float focal, float cw, float ch,
                                 const float* pose, int row0, int row1, const float* rays_in, long n, int ndc,
                                 float near, float far, int S_c, int n_fine, float perturb, float noise,
                                 int white_bkg, unsigned long long seed, long ray0, void* scratch, float* rgb_out,
                                 float* rgb_c_out, void* const* field_events, void* stream) {
    nerf::DeviceGuard device_guard(rgb_out);
    if (pose) n = (long)(row1 - row0) * W;
    if (n < 0 || S_c < 3 || n_fine < 1 || !packed_coarse || !host_tail_coarse || !packed_fine || !host_tail_fine ||
        (n > 0 && (!scratch || !rgb_out || !rgb_c_out || (!pose && !rays_in))))
        return nerf::arg_error("nerf_render_fused");
    if (n == 0) return 0;
    const Scratch L = scratch_layout(n, S_c, n_fine);
    uint8_t* base = static_cast<uint8_t*>(scratch);
    auto f = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
    const int S_f = S_c + n_fine;
    int rc;
    const float* rays = rays_in;
    if (pose) {
        rc = nerf_pack_rays(H, W, focal, cw, ch, pose, row0, row1, nullptr, nullptr, n, ndc, near, far, f(L.rays), stream);
        if (rc) return rc;
        rays = f(L.rays);
    }
    rc = perturb > 0.f ? nerf_sample_coarse_rng(rays, n, S_c, seed, ray0, f(L.z_c), stream)
                       : nerf_sample_coarse(rays, n, S_c, nullptr, f(L.z_c), stream);
    if (rc) return rc;
    // view-direction columns of the packed rays: offset 8, stride 11
    if ((rc = nerf_viewdir_term(packed_coarse, rays + 8, NERF_RAY_STRIDE, 0, n, f(L.vt_c), stream))) return rc;
    if ((rc = nerf_viewdir_term(packed_fine, rays + 8, NERF_RAY_STRIDE, 0, n, f(L.vt_f), stream))) return rc;
    auto mark = [&](int i) {     // optional CUDA events around the two field-kernel launches (bench.py's roofline)
        if (field_events && field_events[i]) cudaEventRecord((cudaEvent_t)field_events[i], (cudaStream_t)stream);
    };
    mark(0);
    if ((rc = nerf_mlp_fwd_host_tail(packed_coarse, host_tail_coarse, NERF_IN_RAYS, rays, f(L.z_c), 0, n * S_c, S_c,
                                     f(L.vt_c), S_c, f(L.raw_c), ray0 * S_c, stream)))
        return rc;
    mark(1);
    rc = noise > 0.f ? nerf_composite_fwd_rng(f(L.raw_c), f(L.z_c), rays + 3, NERF_RAY_STRIDE, noise, seed,
                                              NERF_RNG_STREAM_NOISE_C, ray0, n, S_c, white_bkg, rgb_c_out, f(L.w_c), stream)
                     : nerf_composite_fwd(f(L.raw_c), f(L.z_c), rays + 3, NERF_RAY_STRIDE, nullptr, n, S_c, white_bkg,
                                          rgb_c_out, f(L.w_c), stream);
    if (rc) return rc;
    if ((rc = nerf_resample_merge_rng(f(L.z_c), f(L.w_c), seed, ray0, n, S_c, n_fine, f(L.z_f), stream))) return rc;
    mark(2);
    if ((rc = nerf_mlp_fwd_host_tail(packed_fine, host_tail_fine, NERF_IN_RAYS, rays, f(L.z_f), 0, n * S_f, S_f, f(L.vt_f),
                                     S_f, f(L.raw_f), ray0 * S_f, stream)))
        return rc;
    mark(3);
    return noise > 0.f ? nerf_composite_fwd_rng(f(L.raw_f), f(L.z_f), rays + 3, NERF_RAY_STRIDE, noise, seed,
                                                NERF_RNG_STREAM_NOISE_F, ray0, n, S_f, white_bkg, rgb_out, nullptr, stream)
                       : nerf_composite_fwd(f(L.raw_f), f(L.z_f), rays + 3, NERF_RAY_STRIDE, nullptr, n, S_f, white_bkg,
                                            rgb_out, nullptr, stream);
}
