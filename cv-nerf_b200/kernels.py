"""Tensor-level wrappers over the C ABI (one function per entry point of include/nerf_b200.h).

Each wrapper checks devices/shapes, allocates outputs with torch (device memory + streams are
PyTorch's job here, nothing else) and launches on the tensor's current stream.  Nothing in this
file computes anything on the host.
"""
import ctypes
import dataclasses
import os

import numpy as np
import torch

from . import _lib
from ._lib import f32c, ptr, stream_of

RAY_STRIDE = 11
IN_RAYS, IN_POINTS, IN_EMBEDDED = 0, 1, 2
FLOP_PER_SAMPLE = 2 * 593408   # unpadded MACs of one Model.forward row


# Deterministic accumulation (the default): the gradient kernels write per-CTA partial sums and add them in
# CTA order instead of using floating-point atomics -- two runs give bit-identical gradients and parameters,
# like the reference's CPU autograd.  It costs a second small launch per kernel and no measurable time
# (4.77 ms against 4.84 ms per 4096-ray step on B200, inside box-to-box noise).  set_deterministic(False) or
# NERF_B200_DETERMINISTIC=0 selects the atomic accumulation (red.global.add / atomicAdd) for A/B timing.
DETERMINISTIC = os.environ.get("NERF_B200_DETERMINISTIC", "1") != "0"
_DET_SCRATCH = {}


def set_deterministic(on=True):
    """Switch the gradient kernels to ordered accumulation; returns the previous setting."""
    global DETERMINISTIC
    old, DETERMINISTIC = DETERMINISTIC, bool(on)
    return old


def _det_scratch(kind, nbytes, dev):
    key = (kind, dev)
    buf = _DET_SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _DET_SCRATCH[key] = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
    return buf


class LaunchStats:
    """Counts kernel launches made through this module; when ``timed`` is a list, the field
    kernel additionally brackets itself with CUDA events on its launch stream (bench.py reads
    them after the timed region -- no synchronisation happens here)."""

    def __init__(self):
        self.launches = 0
        self.timed = None

    def reset(self, timed=False):
        self.launches = 0
        self.timed = [] if timed else None


STATS = LaunchStats()


def check(rc, what, launches=1):
    _lib.check(rc, what)
    STATS.launches += launches


def _ndc_consts(height, width, focal):
    """cw, ch folded exactly as the reference's Python expression folds them
    (data_helpers.py:332-333: ``-1./(width/(2.*focal))``), then rounded to fp32."""
    cw = -1. / (width / (2. * focal))
    ch = -1. / (height / (2. * focal))
    return float(np.float32(cw)), float(np.float32(ch))


def _f32(x):
    return float(np.float32(x))


def compute_rays(h, w, f, pose, row0=0, row1=None, want_origins=False):
    """main.py:19-46 -> (origins or None, dirs [(row1-row0), W, 3])."""
    lib = _lib.load()
    pose = f32c(pose[:3, :4])
    row1 = h if row1 is None else row1
    dirs = torch.empty((row1 - row0, w, 3), dtype=torch.float32, device=pose.device)
    origins = torch.empty_like(dirs) if want_origins else None
    check(lib.nerf_compute_rays(h, w, _f32(f), ptr(pose), row0, row1, ptr(origins), ptr(dirs),
                                stream_of(pose)), "nerf_compute_rays")
    return origins, dirs


def get_ndc(height, width, focal, near, o, d):
    lib = _lib.load()
    o, d = f32c(o), f32c(d)
    cw, ch = _ndc_consts(height, width, focal)
    o_out, d_out = torch.empty_like(o), torch.empty_like(d)
    n = o.numel() // 3
    check(lib.nerf_get_ndc(cw, ch, _f32(near), ptr(o), ptr(d), n, ptr(o_out), ptr(d_out), stream_of(o)),
          "nerf_get_ndc")
    return o_out, d_out


def pack_rays(height, width, focal, *, pose=None, row0=0, row1=None, rays_o=None, rays_d=None,
              ndc=True, near=0., far=1.):
    """render() front end (main.py:55-76) -> rays [n,11]."""
    lib = _lib.load()
    cw, ch = _ndc_consts(height, width, focal)
    if pose is not None:
        pose = f32c(pose[:3, :4])
        row1 = height if row1 is None else row1
        n = (row1 - row0) * width
        dev, o, d = pose.device, None, None
    else:
        o, d = f32c(rays_o.reshape(-1, 3)), f32c(rays_d.reshape(-1, 3))
        n, dev, row1 = o.shape[0], o.device, 0
    out = torch.empty((n, RAY_STRIDE), dtype=torch.float32, device=dev)
    check(lib.nerf_pack_rays(height, width, _f32(focal), cw, ch, ptr(pose), row0, row1, ptr(o), ptr(d), n,
                             int(bool(ndc)), _f32(near), _f32(far), ptr(out),
                             torch.cuda.current_stream(dev).cuda_stream), "nerf_pack_rays")
    return out


RNG_T_RAND, RNG_U, RNG_NOISE_C, RNG_NOISE_F = 0, 1, 2, 3      # NERF_RNG_STREAM_* of include/nerf_b200.h


@dataclasses.dataclass(frozen=True)
class Rng:
    """Key of the in-kernel draws of one render_rays call: Philox seed and the global index of the
    batch's first ray (so that chunked / row-sharded calls draw what the whole call would)."""
    seed: int
    ray0: int = 0

    def shifted(self, rays):
        return Rng(self.seed, self.ray0 + int(rays))

    @property
    def key(self):
        return self.seed & 0xFFFFFFFFFFFFFFFF


def rng_fill(kind, rng, stream_id, n, cols, device):
    """The numbers the *_rng kernels draw for rows [rng.ray0, rng.ray0+n) of a stream, as a tensor
    (kind 'uniform' | 'normal')."""
    lib = _lib.load()
    out = torch.empty((n, cols), dtype=torch.float32, device=device)
    check(lib.nerf_rng_fill(0 if kind == "uniform" else 1, rng.key, stream_id, rng.ray0, n, cols, ptr(out),
                            torch.cuda.current_stream(out.device).cuda_stream), "nerf_rng_fill")
    return out


def sample_coarse(rays, n_samples, t_rand=None, rng=None):
    """Coarse depths; stratified jitter from the tensor ``t_rand`` or, with ``rng``, drawn in place."""
    lib = _lib.load()
    n = rays.shape[0]
    z = torch.empty((n, n_samples), dtype=torch.float32, device=rays.device)
    if rng is not None:
        assert t_rand is None
        check(lib.nerf_sample_coarse_rng(ptr(rays), n, n_samples, rng.key, rng.ray0, ptr(z), stream_of(rays)),
              "nerf_sample_coarse_rng")
        return z
    if t_rand is not None:
        t_rand = f32c(t_rand, rays.device)
        assert t_rand.shape == z.shape
    check(lib.nerf_sample_coarse(ptr(rays), n, n_samples, ptr(t_rand), ptr(z), stream_of(rays)),
          "nerf_sample_coarse")
    return z


def _dir_arg(dirs):
    """Accept a [n,3] tensor or a packed [n,11] ray tensor (uses columns 3:6)."""
    if dirs.shape[-1] == RAY_STRIDE:
        return dirs, dirs.data_ptr() + 12, RAY_STRIDE
    dirs = f32c(dirs)
    return dirs, dirs.data_ptr(), 3


@dataclasses.dataclass(frozen=True)
class RngNoise:
    """Density noise drawn inside the compositing kernels: scale * N(0,1) from (rng, stream)."""
    scale: float
    rng: Rng
    stream: int


def composite_fwd(raw, z, dirs, noise=None, white_bkg=False, want_weights=True):
    """noise: None, a [n,S] tensor (already scaled) or an RngNoise."""
    lib = _lib.load()
    raw, z = f32c(raw), f32c(z)
    n, s = z.shape
    keep, dptr, dstride = _dir_arg(dirs)
    rgb = torch.empty((n, 3), dtype=torch.float32, device=raw.device)
    w = torch.empty((n, s), dtype=torch.float32, device=raw.device) if want_weights else None
    if isinstance(noise, RngNoise):
        check(lib.nerf_composite_fwd_rng(ptr(raw), ptr(z), dptr, dstride, float(noise.scale), noise.rng.key, noise.stream,
                                         noise.rng.ray0, n, s, int(bool(white_bkg)), ptr(rgb), ptr(w), stream_of(raw)),
              "nerf_composite_fwd_rng")
        return rgb, w
    noise = None if noise is None else f32c(noise, raw.device)
    check(lib.nerf_composite_fwd(ptr(raw), ptr(z), dptr, dstride, ptr(noise), n, s, int(bool(white_bkg)),
                                 ptr(rgb), ptr(w), stream_of(raw)), "nerf_composite_fwd")
    return rgb, w


def composite_bwd(raw, z, dirs, noise, white_bkg, grad_rgb, grad_w=None):
    lib = _lib.load()
    raw, z = f32c(raw), f32c(z)
    n, s = z.shape
    keep, dptr, dstride = _dir_arg(dirs)
    grad_rgb = f32c(grad_rgb)
    grad_w = None if grad_w is None else f32c(grad_w)
    grad_raw = torch.empty_like(raw)
    if isinstance(noise, RngNoise):
        check(lib.nerf_composite_bwd_rng(ptr(raw), ptr(z), dptr, dstride, float(noise.scale), noise.rng.key, noise.stream,
                                         noise.rng.ray0, n, s, int(bool(white_bkg)), ptr(grad_rgb), ptr(grad_w),
                                         ptr(grad_raw), stream_of(raw)), "nerf_composite_bwd_rng")
        return grad_raw
    noise = None if noise is None else f32c(noise, raw.device)
    check(lib.nerf_composite_bwd(ptr(raw), ptr(z), dptr, dstride, ptr(noise), n, s, int(bool(white_bkg)),
                                 ptr(grad_rgb), ptr(grad_w), ptr(grad_raw), stream_of(raw)),
          "nerf_composite_bwd")
    return grad_raw


def composite_maps(weights, z):
    """weights [n,S], z [n,S] -> [n,3] = (depth, acc, disp)."""
    lib = _lib.load()
    weights, z = f32c(weights), f32c(z)
    n, s = z.shape
    out = torch.empty((n, 3), dtype=torch.float32, device=z.device)
    check(lib.nerf_composite_maps(ptr(weights), ptr(z), n, s, ptr(out), stream_of(z)), "nerf_composite_maps")
    return out


def to_byte(x):
    """uint8(255 * clip(x, 0, 1)) on the device (model.py:134)."""
    lib = _lib.load()
    x = f32c(x)
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(lib.nerf_to_byte(ptr(x), x.numel(), out.data_ptr(), stream_of(x)), "nerf_to_byte")
    return out


def sample_pdf(bins, weights, u):
    lib = _lib.load()
    bins, weights, u = f32c(bins), f32c(weights), f32c(u)
    n, b = bins.shape
    m = u.shape[-1]
    out = torch.empty((n, m), dtype=torch.float32, device=bins.device)
    check(lib.nerf_sample_pdf(ptr(bins), ptr(weights), ptr(u), n, b, m, ptr(out), stream_of(bins)),
          "nerf_sample_pdf")
    return out


def resample_merge(z_c, w_c, u=None, rng=None, n_fine=None):
    """main.py:248-251.  Uniforms from the tensor ``u`` [n,m] or, with ``rng``, drawn in place
    (``n_fine`` = m)."""
    lib = _lib.load()
    z_c, w_c = f32c(z_c), f32c(w_c)
    n, s = z_c.shape
    if rng is not None:
        assert u is None and n_fine is not None
        z_f = torch.empty((n, s + n_fine), dtype=torch.float32, device=z_c.device)
        check(lib.nerf_resample_merge_rng(ptr(z_c), ptr(w_c), rng.key, rng.ray0, n, s, int(n_fine), ptr(z_f),
                                          stream_of(z_c)), "nerf_resample_merge_rng")
        return z_f
    u = f32c(u, z_c.device)
    m = u.shape[-1]
    z_f = torch.empty((n, s + m), dtype=torch.float32, device=z_c.device)
    check(lib.nerf_resample_merge(ptr(z_c), ptr(w_c), ptr(u), n, s, m, ptr(z_f), stream_of(z_c)),
          "nerf_resample_merge")
    return z_f


def render_fused(coarse, fine, *, height, width, focal, pose=None, rows=None, rays=None, ndc=True, near=0., far=1.,
                 n_coarse=64, n_fine=128, perturb=0., noise=0., white_bkg=False, rng=None):
    """The whole render_rays chain (and, with ``pose``, the ray generation of image rows ``rows``) in
    one C call: (rgb [n,3], rgb_c [n,3]).  coarse / fine: model.Model.  Inference only."""
    lib = _lib.load()
    cw, ch = _ndc_consts(height, width, focal)
    if pose is not None:
        pose = f32c(pose[:3, :4])
        r0, r1 = (0, height) if rows is None else rows
        n, dev, rays_ptr = (r1 - r0) * width, pose.device, None
    else:
        rays = f32c(rays)
        r0 = r1 = 0
        n, dev, rays_ptr = rays.shape[0], rays.device, ptr(rays)
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    rgb_c = torch.empty((n, 3), dtype=torch.float32, device=dev)
    if n == 0:
        return rgb, rgb_c
    scratch = torch.empty(int(lib.nerf_render_scratch_bytes(n, n_coarse, n_fine)), dtype=torch.uint8, device=dev)
    pk_c, ht_c, pk_f, ht_f = coarse.packed(), coarse.host_tail(), fine.packed(), fine.host_tail()
    events = None
    if STATS.timed is not None:
        # CUDA events around the two field-kernel launches, recorded by the C entry on the launch stream
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in evs:
            e.record(torch.cuda.current_stream(dev))      # creates the handle; re-recorded inside the call
        events = (ctypes.c_void_p * 4)(*[e.cuda_event for e in evs])
        STATS.timed.append((evs[0], evs[1], n * n_coarse))
        STATS.timed.append((evs[2], evs[3], n * (n_coarse + n_fine)))
    check(lib.nerf_render_fused(pk_c.data_ptr(), ht_c.data_ptr(), pk_f.data_ptr(), ht_f.data_ptr(), height, width,
                                _f32(focal), cw, ch, ptr(pose), r0, r1, rays_ptr, n, int(bool(ndc)), _f32(near), _f32(far),
                                n_coarse, n_fine, float(perturb), float(noise), int(bool(white_bkg)), rng.key, rng.ray0,
                                scratch.data_ptr(), ptr(rgb), ptr(rgb_c), events, torch.cuda.current_stream(dev).cuda_stream),
          "nerf_render_fused", launches=8 + (1 if pose is not None else 0))
    return rgb, rgb_c


def packed_model_bytes():
    return int(_lib.load().nerf_packed_model_bytes())


def pack_model(params, out=None):
    """params: the 24 tensors in registration order (weight, bias per layer) -> uint8 blob."""
    lib = _lib.load()
    params = [f32c(p.detach()) for p in params]
    assert len(params) == 24
    dev = params[0].device
    if out is None:
        out = torch.empty(packed_model_bytes(), dtype=torch.uint8, device=dev)
    arr = (ctypes.c_void_p * 24)(*[ptr(p) for p in params])
    check(lib.nerf_pack_model(arr, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
          "nerf_pack_model", launches=2)
    return out


def viewdir_term(packed, dirs, embedded=False):
    """dirs: [count,3] tensor, packed rays [count,11] (uses columns 8:11) or, with
    embedded=True, rows whose first 27 columns are PE4(dir)."""
    lib = _lib.load()
    count = dirs.shape[0]
    if not embedded and dirs.shape[-1] == RAY_STRIDE:
        dptr, stride = dirs.data_ptr() + 32, RAY_STRIDE
    else:
        dirs = f32c(dirs)
        dptr, stride = dirs.data_ptr(), dirs.shape[-1]
    out = torch.empty((count, 128), dtype=torch.float32, device=dirs.device)
    check(lib.nerf_viewdir_term(packed.data_ptr(), dptr, stride, int(embedded), count, ptr(out),
                                stream_of(dirs)), "nerf_viewdir_term")
    return out


def freq_encode(x, n_freq):
    lib = _lib.load()
    flat = f32c(x.reshape(-1, x.shape[-1]))
    n, dim = flat.shape
    out = torch.empty((n, dim * (1 + 2 * n_freq)), dtype=torch.float32, device=flat.device)
    check(lib.nerf_freq_encode(ptr(flat), n, dim, n_freq, ptr(out), stream_of(flat)), "nerf_freq_encode")
    return out.reshape(*x.shape[:-1], out.shape[-1])


def model_host_tail(packed):
    """Pinned host copy of the blob's biases / l_alpha / l11 for the inference fast path
    (synchronises the current stream once)."""
    lib = _lib.load()
    host = torch.empty(int(lib.nerf_model_host_tail_bytes()), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream(packed.device)
    check(lib.nerf_model_host_tail(packed.data_ptr(), host.data_ptr(), stream.cuda_stream), "nerf_model_host_tail",
          launches=0)
    stream.synchronize()
    return host


def mlp_fwd(packed, mode, in0, in1, rows, samples_per_ray, vterm, vterm_div, in_stride=0,
            probe_layer=None, act_save=None, host_tail=None, row0=0):
    """-> raw [rows,4] (and the probed layer's FP32 activations when probe_layer is given).
    act_save: uint8 buffer of act_bytes(rows) that receives the activation records (training).
    host_tail: model_host_tail(packed) -> the inference fast path (same results).
    row0: global index of the first row when the call is a shard of a larger batch (include/nerf_b200.h)."""
    lib = _lib.load()
    raw = torch.empty((rows, 4), dtype=torch.float32, device=in0.device)
    st = stream_of(in0)
    if probe_layer is None:
        ev = None
        if STATS.timed is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(in0.device))
        if host_tail is not None and act_save is None:
            check(lib.nerf_mlp_fwd_host_tail(packed.data_ptr(), host_tail.data_ptr(), mode, ptr(in0), ptr(in1),
                                             in_stride, rows, samples_per_ray, ptr(vterm), vterm_div, ptr(raw), int(row0), st),
                  "nerf_mlp_fwd_host_tail")
        else:
            check(lib.nerf_mlp_fwd(packed.data_ptr(), mode, ptr(in0), ptr(in1), in_stride, rows, samples_per_ray,
                                   ptr(vterm), vterm_div, ptr(raw),
                                   None if act_save is None else act_save.data_ptr(), int(row0), st), "nerf_mlp_fwd")
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(in0.device))
            STATS.timed.append((ev[0], ev[1], rows))
        return raw
    probe = torch.zeros((rows, 256), dtype=torch.float32, device=in0.device)
    check(lib.nerf_mlp_fwd_probe(packed.data_ptr(), mode, ptr(in0), ptr(in1), in_stride, rows, samples_per_ray,
                                 ptr(vterm), vterm_div, ptr(raw), probe_layer, ptr(probe), st),
          "nerf_mlp_fwd_probe")
    return raw, probe


# ---------------------------------------------------------------------- training (backward)
GRAD_BLOB_FLOATS = None


def act_bytes(rows):
    return int(_lib.load().nerf_mlp_act_bytes(rows))


def dz_bytes(rows):
    return int(_lib.load().nerf_mlp_dz_bytes(rows))


def grad_blob_floats():
    global GRAD_BLOB_FLOATS
    if GRAD_BLOB_FLOATS is None:
        GRAD_BLOB_FLOATS = int(_lib.load().nerf_grad_blob_bytes()) // 4
    return GRAD_BLOB_FLOATS


def pack_model_bwd(params, out=None):
    """Transposed BF16 weights (+ fp32 heads) for the dZ chain."""
    lib = _lib.load()
    params = [f32c(p.detach()) for p in params]
    assert len(params) == 24
    dev = params[0].device
    if out is None:
        out = torch.empty(int(lib.nerf_packed_model_bwd_bytes()), dtype=torch.uint8, device=dev)
    arr = (ctypes.c_void_p * 24)(*[ptr(p) for p in params])
    check(lib.nerf_pack_model_bwd(arr, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
          "nerf_pack_model_bwd", launches=2)
    return out


def pack_models_train(param_lists, packed, packed_bwd):
    """Forward + transposed blobs of one or two Models in a single launch (in place)."""
    lib = _lib.load()
    flat = [f32c(p.detach()) for ps in param_lists for p in ps]
    assert len(flat) == 24 * len(param_lists) and len(packed) == len(packed_bwd) == len(param_lists)
    arr = (ctypes.c_void_p * len(flat))(*[ptr(p) for p in flat])
    out = (ctypes.c_void_p * len(packed))(*[t.data_ptr() for t in packed])
    outb = (ctypes.c_void_p * len(packed_bwd))(*[t.data_ptr() for t in packed_bwd])
    check(lib.nerf_pack_models_train(len(param_lists), arr, out, outb, stream_of(flat[0])), "nerf_pack_models_train")


def mlp_bwd_dz(packed_bwd, grad_raw, act, rows, dz=None, stream=None):
    lib = _lib.load()
    grad_raw = f32c(grad_raw)
    if dz is None:
        dz = torch.empty(dz_bytes(rows), dtype=torch.uint8, device=grad_raw.device)
    check(lib.nerf_mlp_bwd_dz(packed_bwd.data_ptr(), ptr(grad_raw), act.data_ptr(), rows, dz.data_ptr(),
                              _cuda_stream(stream, grad_raw.device)), "nerf_mlp_bwd_dz")
    return dz


def _cuda_stream(stream, dev):
    return (stream if stream is not None else torch.cuda.current_stream(dev)).cuda_stream


def _byte_ptr(t, what):
    """Device pointer of a uint8 record buffer (activation / dZ tile images)."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise _lib.NerfB200Error(f"{what} must be a CUDA tensor; there is no CPU fallback")
    return t.data_ptr()


def mlp_bwd_heads(act, grad_raw, rows, blob, stream=None):
    """dW/db of l_alpha and l11 into the gradient blob (reads the saved h8 / h10 and grad_raw only, so
    it can run as soon as the compositing backward is done)."""
    lib = _lib.load()
    grad_raw = f32c(grad_raw)
    if DETERMINISTIC:
        scratch = _det_scratch(1, lib.nerf_bwd_det_scratch_bytes(1, rows), blob.device)
        check(lib.nerf_mlp_bwd_heads_det(_byte_ptr(act, "act"), ptr(grad_raw), rows, ptr(blob), scratch.data_ptr(),
                                         _cuda_stream(stream, blob.device)), "nerf_mlp_bwd_heads_det", launches=2)
        return blob
    check(lib.nerf_mlp_bwd_heads(_byte_ptr(act, "act"), ptr(grad_raw), rows, ptr(blob), _cuda_stream(stream, blob.device)),
          "nerf_mlp_bwd_heads")
    return blob


def mlp_bwd_dw(act, dz, rows, blob, stream=None):
    """dW/db of the tensor-core layers into the gradient blob (l9/l10: G = dZ10^T h8 into its scratch region)."""
    lib = _lib.load()
    if DETERMINISTIC:
        scratch = _det_scratch(0, lib.nerf_mlp_bwd_dw_det_scratch_bytes(), blob.device)
        check(lib.nerf_mlp_bwd_dw_det(_byte_ptr(act, "act"), _byte_ptr(dz, "dz"), rows, ptr(blob), scratch.data_ptr(),
                                      _cuda_stream(stream, blob.device)), "nerf_mlp_bwd_dw_det", launches=2)
        return blob
    check(lib.nerf_mlp_bwd_dw(_byte_ptr(act, "act"), _byte_ptr(dz, "dz"), rows, ptr(blob),
                              _cuda_stream(stream, blob.device)), "nerf_mlp_bwd_dw")
    return blob


def viewdir_term_bwd(dz, rows, dirs, vterm_div, embedded, blob, stream=None):
    """dW of l10's view-direction columns and db of l10 from the per-ray sums of dZ10."""
    lib = _lib.load()
    if not embedded and dirs.shape[-1] == RAY_STRIDE:
        dptr, stride = dirs.data_ptr() + 32, RAY_STRIDE
    else:
        dirs = f32c(dirs)
        dptr, stride = dirs.data_ptr(), dirs.shape[-1]
    if DETERMINISTIC:
        count = (rows + vterm_div - 1) // vterm_div
        scratch = _det_scratch(2, lib.nerf_bwd_det_scratch_bytes(2, count), blob.device)
        check(lib.nerf_viewdir_term_bwd_det(_byte_ptr(dz, "dz"), dptr, stride, int(embedded), rows, vterm_div, ptr(blob),
                                            scratch.data_ptr(), _cuda_stream(stream, blob.device)),
              "nerf_viewdir_term_bwd_det", launches=2)
        return blob
    check(lib.nerf_viewdir_term_bwd(_byte_ptr(dz, "dz"), dptr, stride, int(embedded), rows, vterm_div, ptr(blob),
                                    _cuda_stream(stream, blob.device)), "nerf_viewdir_term_bwd")
    return blob


def mlp_bwd_unfold(blob, params, stream=None):
    """Gradients of l9 and of l10's first 256 columns from the blob's scratch region (l9 is folded into
    l10, csrc/mlp_layout.h).  params: the Model's 24 tensors in registration order.  Once per backward
    pass, after mlp_bwd_dw and viewdir_term_bwd."""
    lib = _lib.load()
    w9, b9, w10 = (f32c(params[i].detach()) for i in (16, 17, 20))     # l9.weight, l9.bias, l10.weight
    check(lib.nerf_mlp_bwd_unfold(ptr(blob), ptr(w9), ptr(b9), ptr(w10), _cuda_stream(stream, blob.device)),
          "nerf_mlp_bwd_unfold")
    return blob


def mlp_bwd_params(act, dz, grad_raw, rows, dirs, vterm_div, embedded, blob, side_stream=None, params=None):
    """dW/db of every layer accumulated into the fp32 gradient blob (3 launches + the unfold of the
    folded l9/l10 gradients when the Model's ``params`` are given; callers that schedule the launches
    themselves call mlp_bwd_unfold last).

    With ``side_stream`` the two small CUDA-core kernels (l_alpha/l11 heads, l10 view columns) run on
    that stream next to the tensor-core dW kernel instead of after it; the caller's stream waits for
    them before returning.  (train.TrainStep schedules the three launches itself: the heads kernel
    under the dZ chain, the view columns beside dW.)"""
    main = torch.cuda.current_stream(blob.device)
    small = None
    if side_stream is not None:
        side_stream.wait_stream(main)
        small = side_stream
    mlp_bwd_dw(act, dz, rows, blob)
    mlp_bwd_heads(act, grad_raw, rows, blob, stream=small)
    viewdir_term_bwd(dz, rows, dirs, vterm_div, embedded, blob, stream=small)
    if side_stream is not None:
        main.wait_stream(side_stream)
    if params is not None:
        mlp_bwd_unfold(blob, params)
    return blob


def grad_unpack(blob, grads, accumulate=False):
    """Gradient blob -> 24 tensors shaped like Model.parameters() (registration order)."""
    lib = _lib.load()
    assert len(grads) == 24 and all(g.is_contiguous() and g.dtype == torch.float32 for g in grads)
    arr = (ctypes.c_void_p * 24)(*[ptr(g) for g in grads])
    check(lib.nerf_grad_unpack(ptr(blob), arr, int(accumulate), stream_of(blob)), "nerf_grad_unpack")
    return grads


def mse_loss_grad(x, target, want_grad=True, loss=None):
    """mean((x-target)^2) accumulated into `loss` (1-element tensor) and its gradient."""
    lib = _lib.load()
    x, target = f32c(x), f32c(target, x.device)
    assert x.shape == target.shape
    if loss is None:
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    if DETERMINISTIC:
        scratch = _det_scratch(3, lib.nerf_bwd_det_scratch_bytes(3, x.numel()), x.device)
        check(lib.nerf_mse_loss_grad_det(ptr(x), ptr(target), x.numel(), ptr(grad), ptr(loss), scratch.data_ptr(),
                                         stream_of(x)), "nerf_mse_loss_grad_det", launches=2)
        return loss, grad
    check(lib.nerf_mse_loss_grad(ptr(x), ptr(target), x.numel(), ptr(grad), ptr(loss), stream_of(x)),
          "nerf_mse_loss_grad")
    return loss, grad


# ---------------------------------------------------------------------- optimizer / train-loop batch
def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[ptr(t) for t in tensors])


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, betas, eps, step, grad_scale=1.0):
    """torch.optim.Adam semantics for a list of tensors, one launch (in place)."""
    lib = _lib.load()
    n = len(params)
    if n == 0:
        return
    sizes = (ctypes.c_long * n)(*[p.numel() for p in params])
    check(lib.nerf_adam_step(n, _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avg), _ptr_array(exp_avg_sq),
                             sizes, float(lr), float(betas[0]), float(betas[1]), float(eps), int(step),
                             float(grad_scale), stream_of(params[0])), "nerf_adam_step", launches=(n + 47) // 48)


def adam_step_blob(blob, params, exp_avg, exp_avg_sq, lr, betas, eps, step, grad_scale=1.0):
    """Adam for the 24 tensors of one Model with the gradient read from the padded blob."""
    lib = _lib.load()
    assert len(params) == 24
    check(lib.nerf_adam_step_blob(ptr(blob), _ptr_array(params), _ptr_array(exp_avg), _ptr_array(exp_avg_sq),
                                  float(lr), float(betas[0]), float(betas[1]), float(eps), int(step),
                                  float(grad_scale), stream_of(blob)), "nerf_adam_step_blob")


def adam_step_blob_peers(peer_ptrs, params, exp_avg, exp_avg_sq, lr, betas, eps, step, grad_scale, stream):
    """Adam with the gradient summed over the ranks' peer-mapped blobs (device pointers, rank order)."""
    lib = _lib.load()
    assert len(params) == 24
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    check(lib.nerf_adam_step_blob_peers(arr, len(peer_ptrs), _ptr_array(params), _ptr_array(exp_avg),
                                        _ptr_array(exp_avg_sq), float(lr), float(betas[0]), float(betas[1]),
                                        float(eps), int(step), float(grad_scale), stream), "nerf_adam_step_blob_peers")


def train_rays(height, width, focal, pose, n, *, pix=None, seed=0, crop=None, image=None, ndc=True, near=0.,
               far=1., want_pix=False):
    """One train iteration's batch (main.py:351-374): packed rays [n,11], target [n,3] (when an
    image [H,W,3] is given) and optionally the chosen linear pixel indices."""
    lib = _lib.load()
    pose = f32c(pose[:3, :4])
    dev = pose.device
    cw, ch = _ndc_consts(height, width, focal)
    r0, c0, hh, ww = (0, 0, height, width) if crop is None else crop
    if pix is not None:
        pix = pix.to(device=dev, dtype=torch.int32).contiguous()
        n = pix.numel()
    rays = torch.empty((n, RAY_STRIDE), dtype=torch.float32, device=dev)
    target = None
    if image is not None:
        image = f32c(image, dev)
        assert image.shape == (height, width, 3)
        target = torch.empty((n, 3), dtype=torch.float32, device=dev)
    pix_out = torch.empty(n, dtype=torch.int32, device=dev) if want_pix else None
    check(lib.nerf_train_rays(height, width, _f32(focal), cw, ch, ptr(pose), None if pix is None else pix.data_ptr(),
                              int(seed) & 0xFFFFFFFFFFFFFFFF, r0, c0, hh, ww, n, int(bool(ndc)), _f32(near), _f32(far),
                              ptr(image), ptr(rays), ptr(target), None if pix_out is None else pix_out.data_ptr(),
                              torch.cuda.current_stream(dev).cuda_stream), "nerf_train_rays")
    return rays, target, pix_out
