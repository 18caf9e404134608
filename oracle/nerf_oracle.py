"""CPU oracle for the NeRF ray-render hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch restatement, in PyTorch-CPU fp32 arithmetic, of the
algorithm the reference (johnfay11/CV-Nerf) runs for one render / train step.
It exists so that the CUDA path can be checked against something that runs on
a machine without the reference tree.  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the
product package (``cv-nerf_b200/``) never does and has no CPU fallback.

Pinning: every function below is checked bit-for-bit (rays, NDC, sampling) or
to ~1e-6 (anything that goes through MKL sgemm) against outputs of the real
reference functions imported from /root/reference in the build container; the
script ``tests/golden/make_golden.py`` produced the committed fixtures and
``tests/test_oracle_golden.py`` replays them.  The reference itself ships no
tests or golden vectors, so that is the only pin there is.

Reference sites restated (file:line in /root/reference):
  ray_grid            main.py:19-46          (compute_rays)
  ndc_warp            data_helpers.py:327-344 (get_ndc)
  freq_encode         model.py:9-31          (FreqEmbedding.embed)
  field_mlp           model.py:77-107        (Model.forward)
  query_field         model.py:110-131       (net_forward / combine)
  coarse_depths       main.py:221-234
  composite           main.py:170-204        (process_volume_info)
  inverse_cdf_sample  utils.py:4-53          (inv_transform_sampling)
  render_ray_batch    main.py:207-261        (render_rays)
  render_image        main.py:49-99          (render + batch_rays)
  train_loss          main.py:379-383
All random draws are *injected* (see RenderDraws) so that a CUDA run and an
oracle run can consume identical numbers; the reference draws them with
torch.rand / torch.randn in the order documented in SURVEY.md App. A.7.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional, Tuple

import numpy as np
import torch

FAR_DELTA = 1e10          # main.py:175
TRANSMIT_EPS = 1e-10      # main.py:194
PDF_EPS = 1e-5            # utils.py:12

LAYER_NAMES = ("l1", "l2", "l3", "l4", "l5", "l6", "l7", "l8", "l9", "l_alpha", "l10", "l11")
LAYER_SHAPES = {          # (out, in) exactly as nn.Linear stores them, model.py:57-71
    "l1": (256, 63), "l2": (256, 256), "l3": (256, 256), "l4": (256, 256), "l5": (256, 256),
    "l6": (256, 319), "l7": (256, 256), "l8": (256, 256), "l9": (256, 256),
    "l_alpha": (1, 256), "l10": (128, 283), "l11": (3, 128),
}


# --------------------------------------------------------------------------- rays
def ray_grid(h: int, w: int, f, pose: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pinhole rays for an h x w image (main.py:19-46).

    dir = ((j - w/2)/f, -(i - h/2)/f, -1) rotated by pose[:3,:3]; the rotation is
    a broadcast multiply followed by a sum over the last axis, i.e.
    ((d0*R0 + d1*R1) + d2*R2) with separate fp32 roundings.  Origins are the
    translation column broadcast to every pixel.
    """
    pose = torch.as_tensor(pose, dtype=torch.float32)
    col = torch.arange(w, dtype=torch.float32).view(1, w).expand(h, w)   # == linspace(0,w-1,w)
    row = torch.arange(h, dtype=torch.float32).view(h, 1).expand(h, w)
    dx = (col - w * .5) / f                                               # main.py:36
    dy = -(row - h * .5) / f                                              # main.py:37
    cam = torch.stack([dx, dy, -torch.ones_like(dx)], -1)                 # main.py:38
    world = (cam[..., None, :] * pose[:3, :3]).sum(-1)                    # main.py:41-42
    origins = pose[:3, -1].expand(world.shape)                            # main.py:45
    return origins, world


def ndc_warp(height: int, width: int, focal, near: float,
             o: torch.Tensor, d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's NDC re-parameterisation, quirks included (data_helpers.py:327-344).

    Quirk 1: the origin is shifted by t*o (not t*d).  Quirk 2: the direction
    terms use the already-warped origin.  Evaluation order is left to right:
    (cw * ox) / oz.
    """
    t = -(near + o[..., 2]) / d[..., 2]                                   # :329
    o = o + t[..., None] * o                                              # :330
    cw = -1. / (width / (2. * focal))
    ch = -1. / (height / (2. * focal))
    o0 = cw * o[..., 0] / o[..., 2]                                       # :332
    o1 = ch * o[..., 1] / o[..., 2]                                       # :333
    o2 = 1. + 2. * near / o[..., 2]                                       # :334
    o = torch.stack([o0, o1, o2], -1)                                     # :336
    d0 = cw * (d[..., 0] / d[..., 2] - o[..., 0] / o[..., 2])             # :338
    d1 = ch * (d[..., 1] / d[..., 2] - o[..., 1] / o[..., 2])             # :339
    d2 = -2. * near / o[..., 2]                                           # :340
    return o, torch.stack([d0, d1, d2], -1)


def pack_rays(height: int, width: int, focal, rays_o: torch.Tensor, rays_d: torch.Tensor,
              ndc: bool, near: float, far: float) -> torch.Tensor:
    """[N,11] = [o(3) d(3) near far viewdir(3)] (main.py:59-76).  View dirs are
    normalised from the *pre-NDC* directions."""
    view = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    view = view.reshape(-1, 3).float()
    if ndc:
        rays_o, rays_d = ndc_warp(height, width, focal, 1., rays_o, rays_d)
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    ones = torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, near * ones, far * ones, view], -1)


# --------------------------------------------------------------------------- field
def freq_encode(x: torch.Tensor, n_freq: int) -> torch.Tensor:
    """[x, sin(x 2^0), cos(x 2^0), ..., sin(x 2^(L-1)), cos(x 2^(L-1))] (model.py:15-31)."""
    bands = 2. ** torch.linspace(0., n_freq - 1, steps=n_freq)
    parts = [x]
    for b in bands:
        parts.append(torch.sin(x * b))
        parts.append(torch.cos(x * b))
    return torch.cat(parts, -1)


def field_mlp(p: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """8x256 trunk with the skip at layer 6, sigma head, view branch (model.py:77-107).

    ``p`` maps 'l1.weight', 'l1.bias', ... to tensors (a Model.state_dict()).
    Returns [..., 4] = (rgb_raw(3), sigma_raw(1)); no output activations.
    """
    lin = lambda name, v: torch.nn.functional.linear(v, p[name + ".weight"], p[name + ".bias"])
    relu = torch.relu
    enc_pt, enc_dir = x[..., :63], x[..., 63:90]
    h = relu(lin("l1", enc_pt))
    for name in ("l2", "l3", "l4", "l5"):
        h = relu(lin(name, h))
    h = torch.cat([enc_pt, h], -1)                                        # model.py:94 (xyz first)
    for name in ("l6", "l7", "l8"):
        h = relu(lin(name, h))
    sigma = lin("l_alpha", h)                                             # model.py:100
    feat = lin("l9", h)                                                   # model.py:101 (no act)
    h = relu(lin("l10", torch.cat([feat, enc_dir], -1)))
    rgb = lin("l11", h)
    return torch.cat([rgb, sigma], -1)


def query_field(p: Dict[str, torch.Tensor], points: torch.Tensor, viewdirs: torch.Tensor,
                netchunk: Optional[int] = 65536) -> torch.Tensor:
    """points [n,S,3], viewdirs [n,3] -> raw [n,S,4] (model.py:110-131)."""
    n, s, _ = points.shape
    flat = points.reshape(-1, 3)
    enc = freq_encode(flat, 10)
    vd = viewdirs[:, None].expand(points.shape).reshape(-1, 3)
    enc = torch.cat([enc, freq_encode(vd, 4)], -1)
    step = enc.shape[0] if netchunk is None else netchunk
    out = torch.cat([field_mlp(p, enc[i:i + step]) for i in range(0, enc.shape[0], step)], 0)
    return out.reshape(n, s, 4)


# --------------------------------------------------------------------------- sampling / compositing
def coarse_depths(near: torch.Tensor, far: torch.Tensor, n_samples: int,
                  t_rand: Optional[torch.Tensor]) -> torch.Tensor:
    """near/far [n,1] -> z [n,S]; stratified jitter when t_rand [n,S] is given (main.py:221-234)."""
    s = torch.linspace(0., 1., steps=n_samples)
    z = near * (1. - s) + far * s
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = .5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], -1)
        lo = torch.cat([z[..., :1], mid], -1)
        z = lo + (hi - lo) * t_rand
    return z


def composite(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor,
              noise_draw: Optional[torch.Tensor], white_bkg: bool
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """raw [n,S,4], z [n,S], rays_d [n,3] -> rgb_map [n,3], weights [n,S] (main.py:174-204).

    ``noise_draw`` is already scaled (= randn * noise) or None.
    """
    delta = z[..., 1:] - z[..., :-1]
    delta = torch.cat([delta, torch.full_like(delta[..., :1], FAR_DELTA)], -1)
    delta = delta * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] if noise_draw is None else raw[..., 3] + noise_draw
    alpha = 1. - torch.exp(delta * -torch.relu(sigma))                    # main.py:170-171
    trans = torch.cumprod(
        torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + TRANSMIT_EPS], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = (weights[..., None] * rgb).sum(-2)
    if white_bkg:
        rgb_map = rgb_map + (1. - weights.sum(-1)[..., None])
    return rgb_map, weights


def inverse_cdf_sample(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """bins [n,B], weights [n,B-1], u [n,m] in [0,1) -> samples [n,m], unsorted (utils.py:4-53)."""
    w = weights + PDF_EPS
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    lo = (idx - 1).clamp(min=0)
    hi = idx.clamp(max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    span = c_hi - c_lo
    span = torch.where(span < PDF_EPS, torch.ones_like(span), span)
    return (b_hi - b_lo) * ((u - c_lo) / span) + b_lo


@dataclasses.dataclass
class RenderDraws:
    """Random numbers one render_rays call consumes, in the reference's draw order
    (SURVEY.md App. A.7).  None = that draw does not happen."""
    t_rand: Optional[torch.Tensor] = None    # [n,S_c]   uniform, only if perturb > 0
    noise_c: Optional[torch.Tensor] = None   # [n,S_c]   normal * noise, only if noise > 0
    u: Optional[torch.Tensor] = None         # [n,n_fine] uniform, ALWAYS drawn by the reference
    noise_f: Optional[torch.Tensor] = None   # [n,S_c+n_fine] normal * noise

    def rows(self, a: int, b: int) -> "RenderDraws":
        cut = lambda t: None if t is None else t[a:b]
        return RenderDraws(cut(self.t_rand), cut(self.noise_c), cut(self.u), cut(self.noise_f))


def render_ray_batch(rays: torch.Tensor, coarse: Dict[str, torch.Tensor], fine: Dict[str, torch.Tensor],
                     n_coarse: int, n_fine: int, draws: RenderDraws, white_bkg: bool,
                     netchunk: Optional[int] = 65536, extras: bool = False) -> Dict[str, torch.Tensor]:
    """rays [n,11] -> {'rgb_map','rgb_c'} (main.py:207-261)."""
    o, d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    view = rays[:, 8:11]
    z = coarse_depths(near, far, n_coarse, draws.t_rand)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    raw_c = query_field(coarse, pts, view, netchunk)
    rgb_c, w_c = composite(raw_c, z, d, draws.noise_c, white_bkg)
    mids = .5 * (z[..., 1:] + z[..., :-1])
    u = draws.u if draws.u is not None else torch.rand(z.shape[0], n_fine)
    s = inverse_cdf_sample(mids, w_c[..., 1:-1], u).detach()
    z_f, _ = torch.sort(torch.cat([z, s], -1), -1)
    pts_f = o[:, None, :] + d[:, None, :] * z_f[:, :, None]
    raw_f = query_field(fine if fine is not None else coarse, pts_f, view, netchunk)
    rgb_f, w_f = composite(raw_f, z_f, d, draws.noise_f, white_bkg)
    out = {"rgb_map": rgb_f, "rgb_c": rgb_c}
    if extras:
        out.update(z_c=z, raw_c=raw_c, w_c=w_c, z_f=z_f, raw_f=raw_f, w_f=w_f, samples=s)
    return out


def render_image(height: int, width: int, focal, coarse, fine, *, c2w=None, rays=None,
                 ndc: bool = True, near: float = 0., far: float = 1., chunk: int = 32768,
                 n_coarse: int = 64, n_fine: int = 128, draws: Optional[RenderDraws] = None,
                 white_bkg: bool = False, netchunk: Optional[int] = 65536, extras: bool = False):
    """Front end (main.py:49-99): ray generation or a given [2,N,3] batch, view
    dirs, optional NDC, [N,11] packing, chunked render_ray_batch, reshape."""
    if c2w is not None:
        rays_o, rays_d = ray_grid(height, width, focal, c2w)
    else:
        rays_o, rays_d = rays
    lead = list(rays_d.shape[:-1])
    packed = pack_rays(height, width, focal, rays_o, rays_d, ndc, near, far)
    draws = draws or RenderDraws()
    parts = []
    for a in range(0, packed.shape[0], chunk):
        b = min(a + chunk, packed.shape[0])
        parts.append(render_ray_batch(packed[a:b], coarse, fine, n_coarse, n_fine,
                                      draws.rows(a, b), white_bkg, netchunk, extras))
    out = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
    out = {k: v.reshape(lead + list(v.shape[1:])) for k, v in out.items()}
    out["rays"] = packed
    return out


def train_loss(rgb_f: torch.Tensor, rgb_c: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean((rgb-t)^2) + mean((rgb_c-t)^2) (main.py:380-383)."""
    return torch.mean((rgb_f - target) ** 2) + torch.mean((rgb_c - target) ** 2)


# --------------------------------------------------------------------------- helpers for tests / bench
def init_field_params(seed: int, sigma_bias: Optional[float] = None, sigma_gain: float = 1.0
                      ) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Coarse then fine parameter dicts with torch's default nn.Linear init, drawn in
    the reference's construction order (main.py:133-136, model.py:57-71).  With
    ``sigma_bias``/``sigma_gain`` the density head is rescaled so that a random-init
    scene is not degenerate."""
    torch.manual_seed(seed)
    out = []
    for _ in range(2):
        p = {}
        for name in ("l1", "l2", "l3", "l4", "l5", "l6", "l7", "l8", "l9", "l_alpha", "l10", "l11"):
            o, i = LAYER_SHAPES[name]
            lin = torch.nn.Linear(i, o)
            p[name + ".weight"] = lin.weight.detach().clone()
            p[name + ".bias"] = lin.bias.detach().clone()
        if sigma_bias is not None:
            p["l_alpha.bias"].fill_(sigma_bias)
            p["l_alpha.weight"].mul_(sigma_gain)
        out.append(p)
    return out[0], out[1]


def lego_pose(theta: float = -180., phi: float = -30., radius: float = 4.) -> torch.Tensor:
    """Spherical camera pose used for the blender render path (data_helpers.py:34-41)."""
    t = torch.eye(4); t[2, 3] = radius
    ph, th = phi / 180. * np.pi, theta / 180. * np.pi
    rp = torch.tensor([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0],
                       [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0],
                       [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ (rt @ (rp @ t))
