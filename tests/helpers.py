"""Shared helpers for the GPU parity tests (CPU-side emulation of the kernel's rounding points)."""
import os

import numpy as np
import torch

from oracle import nerf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def emulate_field(p, pe63, vterm, probe=None):
    """The kernel's arithmetic restated on the CPU: every tensor-core operand (activations and
    weights) rounded to BF16, FP32+ accumulation, FP32 bias/activation, FP32 sigma and rgb heads.
    pe63 [M,63] fp32 encoding, vterm [M,128] (vterm_reference: includes the folded l9 bias).
    l9 is folded into l10 as the kernel does it (csrc/mlp_layout.h): W' = BF16(W10[:, :256] . W9).
    Returns raw [M,4] (and layer `probe`'s post-activation values: 0..7 = h1..h8, 8 = h10)."""
    W = lambda n: bf16(p[n + ".weight"]).double()
    b = lambda n: p[n + ".bias"].double()
    x = bf16(pe63).double()
    acts = []
    h = torch.relu(x @ W("l1").T + b("l1")); acts.append(h)
    for n in ("l2", "l3", "l4", "l5"):
        h = torch.relu(bf16(h.float()).double() @ W(n).T + b(n)); acts.append(h)
    w6 = W("l6")
    h = torch.relu(x @ w6[:, :63].T + bf16(h.float()).double() @ w6[:, 63:].T + b("l6")); acts.append(h)
    for n in ("l7", "l8"):
        h = torch.relu(bf16(h.float()).double() @ W(n).T + b(n)); acts.append(h)
    sigma = h @ p["l_alpha.weight"].double().T + p["l_alpha.bias"].double()      # fp32 head on fp32 h8
    h10 = torch.relu(bf16(h.float()).double() @ folded_l10_weight(p).T + vterm.double()); acts.append(h10)
    rgb = h10 @ p["l11.weight"].double().T + p["l11.bias"].double()              # fp32 head on fp32 h10
    raw = torch.cat([rgb, sigma], -1).float()
    if probe is None:
        return raw
    return raw, acts[probe].float()


def folded_l10_weight(p):
    """W' = l10.weight[:, :256] . l9.weight, formed in FP32 and rounded to BF16 once (as the pack kernels
    do; model.py:100-104 applies l9 without an activation and feeds only l10) -> fp64 [128,256]."""
    return bf16((p["l10.weight"][:, :256].double() @ p["l9.weight"].double()).float()).double()


def vterm_reference(p, dirs):
    """W10[:,256:283] . PE4(dir) + b10 + W10[:, :256] . b9 (the folded l9 bias) in fp64 -> fp32."""
    pe = O.freq_encode(dirs, 4).double()
    w10 = p["l10.weight"].double()
    return (pe @ w10[:, 256:].T + p["l10.bias"].double() + w10[:, :256] @ p["l9.bias"].double()).float()


def load_model_params(model, p):
    """Copy an oracle parameter dict into a cv_nerf_b200 Model."""
    with torch.no_grad():
        for k, v in p.items():
            name, kind = k.split(".")
            getattr(getattr(model, name), kind).copy_(v)
    return model


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def focal_of(g, prefix=""):
    f = float(g[prefix + "hwf"][2])
    return np.float32(f) if bool(g[prefix + "f_is_f32"]) else f


def flip_aware_stats(got, want, sigma_last_got=None, sigma_last_want=None):
    """max-abs over all rays, and over rays whose far-sample density did not change sign
    (delta_last = 1e10 makes alpha_last a step function of sigma_last)."""
    err = (got - want).abs().amax(-1)
    out = {"max_all": err.max().item(), "n_gt_1e-2": int((err > 1e-2).sum())}
    if sigma_last_got is not None:
        flip = (sigma_last_got > 0) != (sigma_last_want > 0)
        out["n_flip"] = int(flip.sum())
        out["max_noflip"] = err[~flip].max().item() if (~flip).any() else 0.0
    return out


def psnr(a, b):
    mse = torch.mean((a - b) ** 2).item()
    return float("inf") if mse == 0 else -10. * np.log10(mse)


# ------------------------------------------------------------------ training-kernel helpers
def decode_tile_image(buf, n_tiles, tile_bytes, offset, n_blocks):
    """uint8 record buffer -> [n_tiles*128, n_blocks*64] fp32: undo the 128-byte-swizzled BF16 tile
    image layout of csrc/mlp_bwd_layout.h (block = [128 rows][64 cols], 16-byte chunks XOR row&7)."""
    rec = buf.cpu().view(n_tiles, tile_bytes)[:, offset:offset + n_blocks * 16384]
    blocks = rec.reshape(n_tiles, n_blocks, 128, 8, 16)                 # tile, block, row, chunk slot, bytes
    rows = torch.arange(128)
    chunk = torch.arange(8)
    slot = chunk[None, :] ^ (rows[:, None] & 7)                         # [row, logical chunk] -> physical slot
    idx = slot[None, None, :, :, None].expand(n_tiles, n_blocks, 128, 8, 16)
    logical = torch.gather(blocks, 3, idx)                              # logical chunk order
    vals = logical.reshape(n_tiles, n_blocks, 128, 128).contiguous().view(torch.bfloat16)   # 64 bf16 per row
    return vals.permute(0, 2, 1, 3).reshape(n_tiles * 128, n_blocks * 64).float()


def grad_stats(got, want):
    got, want = got.double().reshape(-1), want.double().reshape(-1)
    denom = want.norm().item()
    rel = (got - want).norm().item() / max(denom, 1e-30)
    cos = float(torch.dot(got, want) / max(got.norm().item() * denom, 1e-30))
    return {"rel_l2": rel, "cos": cos, "max_abs": (got - want).abs().max().item(), "ref_norm": denom}


# ------------------------------------------------------------------ measured-parity log
def record(kind, values):
    """Append one measured-parity line to gpurun_out/parity_log.jsonl (copied into profiles/ by hand
    after a GPU run; the file is a by-product, tests never read it)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.environ.get("NERF_B200_PARITY_LOG", os.path.join(root, "gpurun_out", "parity_log.jsonl"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as fh:
            fh.write(json.dumps({"kind": kind, **values}) + "\n")
    except OSError:
        pass
