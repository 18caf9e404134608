"""Drop-in for the render/train call surface of the reference's ``main.py``.

Same names, argument meaning and return structure as /root/reference/main.py:
compute_rays (19-46), render (49-87), batch_rays (90-99; also exported under the name the
reference *calls*, ``batchify_rays``), render_full (102-124), create_model (127-167),
process_volume_info (174-204), render_rays (207-261), decayed_learning_rate (276-277) and a small
reader for the ``configs/*.txt`` flag files (config_parser, 410-457).

All arithmetic runs in libnerf_b200.so.  Random draws (stratified jitter, density noise, the
inverse-CDF uniforms -- SURVEY.md App. A.7) are made INSIDE the consuming kernels from a Philox
counter keyed by (seed, global ray index, sample) -- the seed comes from torch's CPU generator, so
``torch.manual_seed`` makes a run reproducible and no [n,S] tensor of random numbers is ever written
to HBM -- unless a ``RenderDraws`` is passed through the keyword-only ``draws`` argument (parity
tests inject the numbers the oracle consumed).
"""
import dataclasses
import os
import time
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import kernels as K
from ._lib import NerfB200Error, f32c
from .autograd import composite as _composite
from .data_helpers import get_ndc  # noqa: F401  (re-exported like the reference does)
from .model import FreqEmbedding, Model, net_forward, to_byte
from .utils import inv_transform_sampling  # noqa: F401

# rays per internal launch batch; bounds the raw/z scratch (about 6 KB per ray)
MAX_RAYS_PER_LAUNCH = 1 << 20


@dataclasses.dataclass
class RenderDraws:
    """Injected random numbers for one render call, in the reference's draw order."""
    t_rand: Optional[torch.Tensor] = None    # [n,S_c]  uniform (perturb > 0)
    noise_c: Optional[torch.Tensor] = None   # [n,S_c]  normal, unscaled (noise > 0)
    u: Optional[torch.Tensor] = None         # [n,n_fine] uniform (always consumed)
    noise_f: Optional[torch.Tensor] = None   # [n,S_c+n_fine] normal, unscaled

    def rows(self, a, b):
        cut = lambda t: None if t is None else t[a:b]
        return RenderDraws(cut(self.t_rand), cut(self.noise_c), cut(self.u), cut(self.noise_f))


def fresh_rng(ray0=0):
    """A new key for one render call's in-kernel draws, taken from torch's CPU generator (no device
    synchronisation; reproducible under torch.manual_seed)."""
    return K.Rng(int(torch.randint(0, 1 << 62, (1,), dtype=torch.int64).item()), int(ray0))


def compute_rays(h, w, f, pose):
    """Pinhole rays of an h x w image -> (origins [h,w,3] stride-0 view, dirs [h,w,3])."""
    pose = torch.as_tensor(pose)
    if not pose.is_cuda:
        raise NerfB200Error("compute_rays needs the pose on a CUDA device; there is no CPU fallback")
    _, dirs = K.compute_rays(int(h), int(w), f, pose.float())
    origins = pose[:3, -1].float().expand(dirs.shape)
    return origins, dirs


def process_volume_info(raw_rgba, t_samples, r_dirs, noise=0.0, bkg=False, *, noise_draw=None):
    """raw [n,S,4], z [n,S], dirs [n,3] -> (rgb_map [n,3], weights [n,S])."""
    if t_samples.dim() != 2:
        raise NerfB200Error("process_volume_info expects 2-D t_samples [n,S] (as the reference does, main.py:194)")
    nz = None
    if noise > 0.:
        nz = (noise_draw.to(raw_rgba.device) * noise if noise_draw is not None
              else K.RngNoise(float(noise), fresh_rng(), K.RNG_NOISE_C))
    return _composite(raw_rgba, f32c(t_samples), r_dirs, nz, bool(bkg))


def _field(model, rays, z, ray0=0):
    n, s = z.shape
    spec = dict(mode=K.IN_RAYS, in0=rays, in1=z, rows=n * s, samples=s, vterm_div=s, dirs=rays, row0=int(ray0) * s)
    return model.field(spec).reshape(n, s, 4)


def render_rays(ray_batch, coarse_model, q_fn=None, n_coarse_samples=64, perturb=0.0, n_fine_samples=0,
                fine_model=None, white_bkg=False, noise=0.0, *, draws=None, rng=None, extras=False, maps=False):
    """[n,11] rays -> {'rgb_map': [n,3], 'rgb_c': [n,3]} (main.py:207-261).

    ``q_fn`` is accepted and ignored: when the models are cv_nerf_b200 ``Model`` instances the
    encode+MLP closure the reference builds (main.py:138-141) is what the fused kernel does.
    ``draws`` injects random numbers; whatever it leaves out is drawn in-kernel from ``rng``
    (a kernels.Rng; default: a fresh key from torch's CPU generator)."""
    if not isinstance(coarse_model, Model) or (fine_model is not None and not isinstance(fine_model, Model)):
        raise NerfB200Error("render_rays needs cv_nerf_b200.model.Model networks; there is no fallback path")
    if ray_batch.shape[-1] != K.RAY_STRIDE:
        raise NerfB200Error("render_rays expects packed rays [n,11] (o,d,near,far,viewdir)")
    rays = f32c(ray_batch)
    if not rays.is_cuda:
        raise NerfB200Error("render_rays needs CUDA tensors; there is no CPU fallback")
    n, dev = rays.shape[0], rays.device
    draws = draws or RenderDraws()
    rng = rng or fresh_rng()
    on_dev = lambda t: None if t is None else f32c(t, dev)

    if not perturb > 0.:
        z_c = K.sample_coarse(rays, n_coarse_samples)
    elif draws.t_rand is not None:
        z_c = K.sample_coarse(rays, n_coarse_samples, on_dev(draws.t_rand))
    else:
        z_c = K.sample_coarse(rays, n_coarse_samples, rng=rng)

    def noise_for(injected, stream):
        if not noise > 0.:
            return None
        return on_dev(injected) * noise if injected is not None else K.RngNoise(float(noise), rng, stream)

    want_all = extras or maps
    raw_c = _field(coarse_model, rays, z_c, rng.ray0)
    rgb_c, w_c = _composite(raw_c, z_c, rays, noise_for(draws.noise_c, K.RNG_NOISE_C), bool(white_bkg))

    if draws.u is not None:
        z_f = K.resample_merge(z_c, w_c.detach(), on_dev(draws.u))
    else:
        z_f = K.resample_merge(z_c, w_c.detach(), rng=rng, n_fine=n_fine_samples)

    run = coarse_model if fine_model is None else fine_model
    raw_f = _field(run, rays, z_f, rng.ray0)
    # the fine pass's weights are only materialised when somebody reads them (0.49 GB per 800x800 frame)
    rgb_f, w_f = _composite(raw_f, z_f, rays, noise_for(draws.noise_f, K.RNG_NOISE_F), bool(white_bkg),
                            want_weights=want_all)

    out = {'rgb_map': rgb_f, 'rgb_c': rgb_c}
    if maps:
        # not returned by the reference (main.py:259-261); depth/acc/disp of the fine pass
        m = K.composite_maps(w_f.detach(), z_f)
        out.update(depth_map=m[:, 0], acc_map=m[:, 1], disp_map=m[:, 2])
    if extras:
        out.update(z_c=z_c, raw_c=raw_c, w_c=w_c, z_f=z_f, raw_f=raw_f, w_f=w_f)
    return out


def batch_rays(rays_flat, chunk=32768, *, draws=None, rng=None, **kwargs):
    """Chunked render_rays (main.py:90-99).  ``chunk`` exists in the reference to bound the
    [chunk*S,90] encoding tensor; the inference path has no such tensor (about 6 KB of raw/depth
    scratch per ray), so without autograd chunks are merged up to MAX_RAYS_PER_LAUNCH rays per
    launch.  With autograd recording, the field saves about 1.35 MB of activations per ray
    (675 840 B per 128-sample tile x 256 samples), so there ``chunk`` is honoured as the memory
    bound it is in the reference."""
    chunk = max(int(chunk), 1)
    step = chunk if torch.is_grad_enabled() else max(chunk, MAX_RAYS_PER_LAUNCH)
    res = {}
    rng = rng or fresh_rng()
    for i in range(0, rays_flat.shape[0], step):
        d = None if draws is None else draws.rows(i, i + step)
        ret = render_rays(rays_flat[i:i + step], draws=d, rng=rng.shifted(i), **kwargs)
        for k in ret:
            res.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in res.items()}


batchify_rays = batch_rays  # the name render() calls in the reference (main.py:79)
to8b = to_byte              # the name main() calls in the reference (main.py:404)


def _fused_inference(kwargs, draws):
    """True when a render() call can take the one-call whole-chain entry (nerf_render_fused):
    inference (no autograd recording), in-kernel random draws, both networks given, nothing but the
    two colour maps asked for."""
    coarse, fine = kwargs.get('coarse_model'), kwargs.get('fine_model')
    if draws is not None or kwargs.get('extras') or kwargs.get('maps'):
        return False
    if not isinstance(coarse, Model) or not isinstance(fine, Model) or not kwargs.get('n_fine_samples', 0) > 0:
        return False
    if kwargs.get('n_coarse_samples', 64) < 3:
        return False
    if torch.is_grad_enabled() and any(p.requires_grad for m in (coarse, fine) for p in m.parameters()):
        return False
    return os.environ.get("NERF_B200_FUSED_RENDER", "1") != "0"


def _render_fused(height, width, focal, rays, c2w, ndc, near, far, rows, rng, *, coarse_model, fine_model,
                  n_coarse_samples=64, n_fine_samples=0, perturb=0.0, white_bkg=False, noise=0.0, q_fn=None, **unused):
    """render() for inference through nerf_render_fused: one C call per batch of up to
    MAX_RAYS_PER_LAUNCH rays sequences every launch of the chain."""
    common = dict(height=height, width=width, focal=focal, ndc=ndc, near=near, far=far, n_coarse=int(n_coarse_samples),
                  n_fine=int(n_fine_samples), perturb=float(perturb), noise=float(noise), white_bkg=bool(white_bkg))
    parts = []
    if c2w is not None:
        c2w = torch.as_tensor(c2w)
        if not c2w.is_cuda:
            raise NerfB200Error("render needs c2w on a CUDA device; there is no CPU fallback")
        r0, r1 = (0, height) if rows is None else rows
        rng = fresh_rng(0) if rng is None else rng
        step = max(MAX_RAYS_PER_LAUNCH // width, 1)
        for a in range(r0, r1, step):
            b = min(a + step, r1)
            parts.append(K.render_fused(coarse_model, fine_model, pose=c2w.float(), rows=(a, b),
                                        rng=rng.shifted(a * width), **common))
        lead = [r1 - r0, width]
    else:
        rays_o, rays_d = rays
        if not rays_d.is_cuda:
            raise NerfB200Error("render needs CUDA ray tensors; there is no CPU fallback")
        packed = K.pack_rays(height, width, focal, rays_o=rays_o, rays_d=rays_d, ndc=ndc, near=near, far=far)
        rng = fresh_rng(0) if rng is None else rng
        for a in range(0, packed.shape[0], MAX_RAYS_PER_LAUNCH):
            parts.append(K.render_fused(coarse_model, fine_model, rays=packed[a:a + MAX_RAYS_PER_LAUNCH],
                                        rng=rng.shifted(a), **common))
        lead = list(rays_d.shape[:-1])
        if not parts:
            empty = torch.empty((0, 3), dtype=torch.float32, device=rays_d.device)
            parts.append((empty, empty.clone()))
    rgb = parts[0][0] if len(parts) == 1 else torch.cat([p[0] for p in parts], 0)
    rgb_c = parts[0][1] if len(parts) == 1 else torch.cat([p[1] for p in parts], 0)
    return [rgb.reshape(lead + [3]), {'rgb_c': rgb_c.reshape(lead + [3])}]


def render(height, width, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1.,
           *, rows=None, draws=None, rng=None, **kwargs):
    """Full front end (main.py:49-87): returns ``[rgb_map, {'rgb_c': ...}]`` shaped like the
    ray batch (``[H,W,3]`` for ``c2w``).  ``rows=(r0,r1)`` restricts a ``c2w`` render to image
    rows [r0,r1) (used to shard a frame across GPUs; the in-kernel draws are keyed by the global
    ray index, so the shards of a frame rendered with the same ``rng`` seed equal the whole frame)."""
    height, width = int(height), int(width)
    first_ray = 0
    fused = _fused_inference(kwargs, draws)
    if fused:
        return _render_fused(height, width, focal, rays, c2w, ndc, near, far, rows, rng, **kwargs)
    if c2w is not None:
        c2w = torch.as_tensor(c2w)
        if not c2w.is_cuda:
            raise NerfB200Error("render needs c2w on a CUDA device; there is no CPU fallback")
        r0, r1 = (0, height) if rows is None else rows
        packed = K.pack_rays(height, width, focal, pose=c2w.float(), row0=r0, row1=r1, ndc=ndc, near=near, far=far)
        lead = [r1 - r0, width]
        first_ray = r0 * width
    else:
        rays_o, rays_d = rays
        if not rays_d.is_cuda:
            raise NerfB200Error("render needs CUDA ray tensors; there is no CPU fallback")
        packed = K.pack_rays(height, width, focal, rays_o=rays_o, rays_d=rays_d, ndc=ndc, near=near, far=far)
        lead = list(rays_d.shape[:-1])
    rng = fresh_rng(first_ray) if rng is None else rng.shifted(first_ray)
    all_ret = batchify_rays(packed, chunk, draws=draws, rng=rng, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], lead + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


def render_full(render_poses, hwf, chunk, render_kwargs, save_dir=None, factor=0, *, as_bytes=False,
                process_group=None, verbose=True):
    """Render every pose (main.py:102-124) -> np.ndarray [n_poses,H,W,3] (float32, or uint8 with
    ``as_bytes`` -- the to_byte quantisation of main.py:118 done on the device, 4x less to copy).

    Frames are copied to pinned host memory asynchronously, so frame i+1 renders while frame i
    drains (the reference synchronises on ``.cpu().numpy()`` per frame, main.py:116).  When
    torch.distributed is initialised the frames are rendered round-robin across the ranks
    (frame-parallel: no communication until the final gather) and every rank returns the full stack."""
    from . import parallel as P
    height, width, focal = hwf
    if factor != 0:
        height, width, focal = height // factor, width // factor, focal / factor
    height, width = int(height), int(width)
    n = len(render_poses)
    rank, world = P.world_info(process_group)
    mine = P.frame_indices(n, world, rank)
    dtype = torch.uint8 if as_bytes else torch.float32
    t = time.time()
    if world == 1:
        host = torch.empty((n, height, width, 3), dtype=dtype).pin_memory()
        with torch.no_grad():
            for i, c2w in enumerate(render_poses):
                rgb, _ = render(height, width, focal, chunk=chunk, c2w=torch.as_tensor(c2w)[:3, :4], **render_kwargs)
                host[i].copy_(K.to_byte(rgb) if as_bytes else rgb, non_blocking=True)
        torch.cuda.synchronize()
        rgbs = host.numpy()
    else:
        with torch.no_grad():
            frames = []
            for i in mine:
                rgb, _ = render(height, width, focal, chunk=chunk, c2w=torch.as_tensor(render_poses[i])[:3, :4],
                                **render_kwargs)
                frames.append(K.to_byte(rgb) if as_bytes else rgb)
            dev = torch.device("cuda", torch.cuda.current_device())
            local = torch.stack(frames, 0) if frames else torch.empty((0, height, width, 3), dtype=dtype, device=dev)
            rgbs = P.gather_frames(local, n, process_group).cpu().numpy()
    if save_dir is not None and rank == 0:
        os.makedirs(save_dir, exist_ok=True)
        for i in range(n):      # main.py:117-120 writes '{:03d}.png' of the 8-bit frame
            write_png(os.path.join(save_dir, '{:03d}.png'.format(i)), rgbs[i] if as_bytes else to_byte(rgbs[i]))
    if verbose and rank == 0:
        print(f"rendered {n} frames in {time.time() - t:.3f}s")
    return rgbs


def write_png(path, rgb8):
    """[H,W,3] uint8 -> 8-bit RGB PNG (the reference uses imageio.imwrite, main.py:120; imageio is
    not a dependency here, and a PNG is two zlib calls)."""
    import struct
    import zlib
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    rows = np.concatenate([np.zeros((h, 1), np.uint8), rgb8.reshape(h, w * 3)], 1).tobytes()   # filter 0 per row

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xffffffff)
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                 + chunk(b"IDAT", zlib.compress(rows, 6)) + chunk(b"IEND", b""))


def create_model(args):
    """Embedders, coarse+fine Model, q_fn closure, Adam and the two kwargs dicts
    (main.py:127-167).  Construction order (coarse, then fine) matches the reference so a given
    torch seed produces the same weights."""
    device = torch.device("cuda")
    xyz_embedder = FreqEmbedding(10)
    ang_embedder = FreqEmbedding(4)
    coarse_model = Model().to(device)
    grad_vars = list(coarse_model.parameters())
    fine_model = Model().to(device)
    grad_vars += list(fine_model.parameters())

    q_fn = lambda inputs, dirs, network_fn: net_forward(inputs, dirs, network_fn,
                                                        embed_fn=xyz_embedder.embed,
                                                        embeddirs_fn=ang_embedder.embed,
                                                        netchunk=args.netchunk)
    # torch.optim.Adam's update rule as one multi-tensor launch (train.FusedAdam)
    from .train import FusedAdam
    optimizer = FusedAdam(params=grad_vars, lr=args.lr, betas=(0.9, 0.999))
    start = 0
    render_kwargs_train = {
        'q_fn': q_fn,
        'perturb': args.perturb,
        'n_fine_samples': args.n_fine_samples,
        'fine_model': fine_model,
        'n_coarse_samples': args.n_coarse_samples,
        'coarse_model': coarse_model,
        'white_bkg': args.white_bkg,
        'noise': args.noise,
    }
    if args.dtype != 'llff' or args.no_ndc:
        render_kwargs_train['ndc'] = False
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['noise'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


def decayed_learning_rate(step, decay_steps, initial_lr, decay_rate=0.1):
    return initial_lr * (decay_rate ** (step / decay_steps))


# ------------------------------------------------------------------ configs/*.txt reader
_FLAG_DEFAULTS = dict(  # config_parser defaults, main.py:410-457
    name=None, base_dir='./logs/', data_dir='./data/llff/fern', save_dir='./logs', n_rays=4096, lr=5e-4,
    lr_decay=250, chunk=1024 * 32, netchunk=1024 * 64, no_reload=False, ft_path=None, n_coarse_samples=64,
    n_fine_samples=0, perturb=1., noise=0., render_only=False, render_test=False, render_factor=0,
    precrop_iters=0, precrop_frac=None, testskip=8, white_bkg=False, half_res=False, factor=8, no_ndc=False,
    spherify=False, llffhold=8, print_freq=100, vid_freq=5000, dtype='llff')


def load_config(path=None, **overrides):
    """Parse a ``key = value`` flag file (configs/lego.txt, configs/fern.txt, ...) into a namespace
    with the reference's defaults."""
    flags = dict(_FLAG_DEFAULTS)
    if path is not None:
        with open(path) as fh:
            for line in fh:
                line = line.split('#', 1)[0].strip()
                if not line or '=' not in line:
                    continue
                key, val = (s.strip() for s in line.split('=', 1))
                if key not in flags:
                    continue  # unknown keys are ignored (configs/skull-orig.txt has leftovers)
                proto = _FLAG_DEFAULTS[key]
                if isinstance(proto, bool):
                    flags[key] = val.lower() in ('true', '1', 'yes')
                elif isinstance(proto, int):
                    flags[key] = int(val)
                elif isinstance(proto, float) or key == 'precrop_frac':
                    flags[key] = float(val)
                else:
                    flags[key] = val
    flags.update(overrides)
    return SimpleNamespace(**flags)
