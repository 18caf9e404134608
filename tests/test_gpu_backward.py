"""GPU parity of the training path: saved activations, dZ chain, dW/db contractions, and the whole
loss.backward() of a render against the reference's autograd (oracle + fixtures recorded from the
real reference: gradient norms and the first 96 elements of every parameter gradient).

Tolerance: the contractions run in BF16 (operands) with FP32 accumulation, so gradients are compared
by relative L2 error and cosine per tensor (SURVEY.md App. C measured rel-L2 1.2e-2 for a
BF16-emulated reference); the bounds below are written per test."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.helpers import (bf16, decode_tile_image, focal_of, folded_l10_weight, golden, grad_stats, load_model_params,
                           record)

pytestmark = pytest.mark.gpu
DEV = "cuda"
ACT_TILE, DZ_TILE = 610304, 557056     # csrc/mlp_bwd_layout.h: kActTileBytes, kDzTileBytes (no h9 / dZ9: l9 is folded into l10)


def _K():
    import cv_nerf_b200
    return cv_nerf_b200.kernels


def _model(p):
    from cv_nerf_b200.model import Model
    return load_model_params(Model(), p).to(DEV)


def _reference_chain(p, pts, dirs_per_row, grad_raw):
    """fp32 autograd reference of the field with every layer's pre-activation kept:
    returns (raw, acts dict, dz dict, param grads dict)."""
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    x = torch.cat([O.freq_encode(pts, 10), O.freq_encode(dirs_per_row, 4)], -1)
    lin = lambda n, v: torch.nn.functional.linear(v, q[n + ".weight"], q[n + ".bias"])
    pe, ped = x[:, :63], x[:, 63:]
    z, acts = {}, {}
    h = pe
    for i, n in enumerate(("l1", "l2", "l3", "l4", "l5"), 1):
        z[i] = lin(n, h); h = torch.relu(z[i]); acts[i] = h
    h = torch.cat([pe, h], -1)
    for i, n in ((6, "l6"), (7, "l7"), (8, "l8")):
        z[i] = lin(n, h); h = torch.relu(z[i]); acts[i] = h
    sigma = lin("l_alpha", h)
    z[9] = lin("l9", h); acts[9] = z[9]
    z[10] = lin("l10", torch.cat([z[9], ped], -1)); acts[10] = torch.relu(z[10])
    rgb = lin("l11", acts[10])
    raw = torch.cat([rgb, sigma], -1)
    for v in z.values():
        v.retain_grad()
    raw.backward(grad_raw)
    return raw.detach(), {k: v.detach() for k, v in acts.items()}, {k: v.grad for k, v in z.items()}, \
        {k: v.grad for k, v in q.items()}


def _emulated_chain(p, pe, acts, pe_dir, grad_raw, mask_acts=None):
    """The kernels' arithmetic restated in fp64: BF16 operands (saved activations, BF16-rounded dZ,
    BF16 weights), exact accumulation, ReLU masks taken from the SAVED activations -- or, with
    ``mask_acts``, from another forward's activations (the fp32 reference's: isolates the rounding
    error of the contractions from the ReLU-mask flips of a BF16 forward).
    Returns (dz dict, param grads dict)."""
    W = lambda n: bf16(p[n + ".weight"]).double()
    g = grad_raw.double()
    vals = {k: v.double() for k, v in acts.items()}
    a = vals if mask_acts is None else {k: v.double() for k, v in mask_acts.items()}
    dz, r = {}, {}
    r16 = lambda t: bf16(t.float()).double()
    dz[10] = r16((g[:, :3] @ p["l11.weight"].double()) * (a[10] > 0))
    # l9 folded into l10: dZ8 comes straight from dZ10 through W' = BF16(W10[:, :256] . W9); there is no dZ9
    dz[8] = r16((dz[10] @ folded_l10_weight(p) + g[:, 3:4] * p["l_alpha.weight"].double()) * (a[8] > 0))
    dz[7] = r16((dz[8] @ W("l8")) * (a[7] > 0))
    dz[6] = r16((dz[7] @ W("l7")) * (a[6] > 0))
    dz[5] = r16((dz[6] @ W("l6")[:, 63:]) * (a[5] > 0))
    for i in (4, 3, 2, 1):
        dz[i] = r16((dz[i + 1] @ W(f"l{i + 1}")) * (a[i] > 0))
    x_in = {1: pe.double()[:, :63], 6: torch.cat([pe.double()[:, :63], vals[5]], -1)}
    for i in range(1, 9):
        x = x_in.get(i, vals.get(i - 1))
        r[f"l{i}.weight"] = dz[i].T @ x
        r[f"l{i}.bias"] = dz[i].sum(0)
    dv = dz[10]
    # the unfold (nerf_mlp_bwd_unfold): G = dZ10^T h8, then the chain rule through feat = W9 h8 + b9 in FP32
    G, db10 = dv.T @ vals[8], dv.sum(0)
    w9, b9, w10a = p["l9.weight"].double(), p["l9.bias"].double(), p["l10.weight"][:, :256].double()
    r["l10.weight"] = torch.cat([G @ w9.T + db10[:, None] * b9[None, :], dv.T @ pe_dir.double()], -1)
    r["l10.bias"] = db10
    r["l9.weight"] = w10a.T @ G
    r["l9.bias"] = w10a.T @ db10
    r["l_alpha.weight"] = (g[:, 3:4] * vals[8]).sum(0, keepdim=True)
    r["l_alpha.bias"] = g[:, 3].sum().reshape(1)
    r["l11.weight"] = g[:, :3].T @ vals[10]
    r["l11.bias"] = g[:, :3].sum(0)
    return dz, r


def _fp32_bar(layer):
    """rel-L2 bound of a BF16-forward gradient against the fp32 reference, by depth of the tensor in
    the backward chain.  A ReLU mask that differs between the two forwards moves that element by its
    full size, so the error is ~sqrt(fraction of masks flipped) per layer (measured 4e-4 -> 0.02 at
    dZ10) and accumulates towards l1: measured on B200 0.020 (dZ10) ... 0.119 (dZ1), 0.133 (l1.bias)
    (profiles/r02_grad_parity.txt); the bounds are the measured worst cases plus about a quarter."""
    return {10: 0.030, 9: 0.030, 8: 0.060, 7: 0.080, 6: 0.095, 5: 0.105, 4: 0.115, 3: 0.130, 2: 0.145, 1: 0.155}[layer]


@pytest.mark.parametrize("rows,S", [(96, 8), (1000, 8), (3 * 128 * 4, 64)])
def test_field_backward_stages(rows, S):
    """saved activations, dZ of every layer and all 24 parameter gradients, stage by stage."""
    K = _K()
    torch.manual_seed(rows)
    p, _ = O.init_field_params(3, 0.5, 5.0)
    model = _model(p)
    n_rays = rows // S
    pts = (torch.rand(rows, 3) * 2 - 1) * 1.5
    dirs = torch.nn.functional.normalize(torch.randn(n_rays, 3), dim=-1)
    grad_raw = torch.randn(rows, 4) * torch.tensor([1., 1., 1., .3])
    raw_ref, acts, dz_ref, g_ref = _reference_chain(p, pts, dirs.repeat_interleave(S, 0), grad_raw)

    packed = model.packed()
    vterm = K.viewdir_term(packed, dirs.to(DEV))
    act = torch.zeros(K.act_bytes(rows), dtype=torch.uint8, device=DEV)
    raw = K.mlp_fwd(packed, K.IN_POINTS, pts.to(DEV), None, rows, S, vterm, S, act_save=act)
    raw_plain = K.mlp_fwd(packed, K.IN_POINTS, pts.to(DEV), None, rows, S, vterm, S)
    torch.cuda.synchronize()
    assert torch.equal(raw, raw_plain), "saving activations changed the forward result"
    n_tiles = (rows + 127) // 128
    assert act.numel() == n_tiles * ACT_TILE

    # 1. activation records
    pe = decode_tile_image(act, n_tiles, ACT_TILE, 0, 1)[:rows]
    want_pe = torch.cat([O.freq_encode(pts, 10), torch.zeros(rows, 1)], -1)
    assert (pe - bf16(want_pe)).abs().max() <= 2e-2          # bf16 ulp at |x|<=2 is 1.6e-2 (sin/cos anchors)
    saved = {}
    for i in range(1, 9):
        saved[i] = decode_tile_image(act, n_tiles, ACT_TILE, 16384 + (i - 1) * 65536, 4)[:rows]
        err = (saved[i] - acts[i]).abs().max().item()
        assert err <= 2e-2 * max(1.0, acts[i].abs().max().item()), (i, err)
    saved[10] = decode_tile_image(act, n_tiles, ACT_TILE, 16384 + 8 * 65536, 2)[:rows]
    assert (saved[10] - acts[10]).abs().max() <= 2e-2 * max(1.0, acts[10].abs().max().item())
    pe_dir = O.freq_encode(dirs.repeat_interleave(S, 0), 4)
    dz_emu, g_emu = _emulated_chain(p, pe, saved, pe_dir, grad_raw)
    # the same BF16 arithmetic with the fp32 forward's ReLU masks: what is left against the fp32
    # reference is the rounding of the contractions alone
    dz_fm, g_fm = _emulated_chain(p, pe, saved, pe_dir, grad_raw, mask_acts=acts)
    flipped = {i: ((saved[i] > 0) != (acts[i] > 0)).float().mean().item() for i in (1, 2, 3, 4, 5, 6, 7, 8, 10)}
    print("fraction of ReLU masks that differ between the BF16 and the fp32 forward:", flipped)
    record("relu_mask_flips", dict(rows=rows, **{f"l{i}": v for i, v in flipped.items()}))

    # 2. dZ chain
    dz = K.mlp_bwd_dz(model.packed_bwd(), grad_raw.to(DEV), act, rows)
    torch.cuda.synchronize()
    assert dz.numel() == n_tiles * DZ_TILE
    full = decode_tile_image(dz, n_tiles, DZ_TILE, 8 * 65536, 2)
    assert full[rows:].abs().max().item() == 0 if full.shape[0] > rows else True, "padding rows must carry zero gradient"
    for i in (10, 8, 7, 6, 5, 4, 3, 2, 1):          # no dZ9: l9 is folded into l10
        off, nb = (8 * 65536, 2) if i == 10 else ((i - 1) * 65536, 4)
        got = decode_tile_image(dz, n_tiles, DZ_TILE, off, nb)[:rows]
        # (a) the kernel's own arithmetic (BF16 operands, masks from the saved activations): tight
        se = grad_stats(got, dz_emu[i])
        # (b) the fp32 reference: ReLU masks of near-zero activations differ between a BF16 and an
        #     fp32 forward, which moves single elements by O(|dh|); judged by norm and direction
        st = grad_stats(got, dz_ref[i])
        print(f"dZ{i} vs emulation", se, "vs fp32 reference", st)
        record("grad_stage", dict(rows=rows, tensor=f"dZ{i}", emu_rel_l2=se["rel_l2"], emu_cos=se["cos"],
                                  fp32_rel_l2=st["rel_l2"], fp32_cos=st["cos"]))
        sm = grad_stats(dz_fm[i], dz_ref[i])
        record("grad_stage_fp32_masks", dict(rows=rows, tensor=f"dZ{i}", rel_l2=sm["rel_l2"], cos=sm["cos"]))
        assert se["rel_l2"] <= 2e-3 and se["cos"] >= 0.99999, (i, se)          # measured <= 5.7e-4
        assert sm["rel_l2"] <= 3e-2, (i, sm)                                     # rounding alone
        assert st["rel_l2"] <= _fp32_bar(i) and st["cos"] >= 0.99, (i, st)

    # 3. parameter gradients
    blob = torch.zeros(K.grad_blob_floats(), device=DEV)
    K.mlp_bwd_params(act, dz, grad_raw.to(DEV), rows, dirs.to(DEV), S, False, blob, params=model.ordered_params())
    grads = [torch.empty_like(q) for q in model.ordered_params()]
    K.grad_unpack(blob, grads)
    torch.cuda.synchronize()
    names = [f"{n}.{k}" for n in O.LAYER_NAMES for k in ("weight", "bias")]
    for name, got in zip(names, grads):
        se = grad_stats(got.cpu(), g_emu[name])
        st = grad_stats(got.cpu(), g_ref[name])
        print(name, "vs emulation", se, "vs fp32 reference", st)
        record("grad_stage", dict(rows=rows, tensor=name, emu_rel_l2=se["rel_l2"], emu_cos=se["cos"],
                                  fp32_rel_l2=st["rel_l2"], fp32_cos=st["cos"]))
        sm = grad_stats(g_fm[name], g_ref[name])
        record("grad_stage_fp32_masks", dict(rows=rows, tensor=name, rel_l2=sm["rel_l2"], cos=sm["cos"]))
        layer = {"l_alpha": 9, "l10": 10, "l11": 11}.get(name.split(".")[0]) or int(name.split(".")[0][1:])
        assert se["rel_l2"] <= 2e-3 and se["cos"] >= 0.99999, (name, se)
        assert sm["rel_l2"] <= 3e-2, (name, sm)
        # parameter gradients are sums of dZ over the rows (biases: plain sums), the flips do not
        # average out the same way: measured up to 1.37x the dZ figure of the layer
        assert st["rel_l2"] <= _fp32_bar(min(layer, 10)) * 1.4 and st["cos"] >= 0.99, (name, st)
    # accumulate mode adds
    K.grad_unpack(blob, grads, accumulate=True)
    torch.cuda.synchronize()
    st = grad_stats(grads[2].cpu(), 2 * g_emu["l2.weight"])
    assert st["rel_l2"] <= 1e-2


def test_model_forward_backward_through_autograd():
    """Model.forward(x[...,90]) is differentiable w.r.t. the parameters (embedded-input mode)."""
    K = _K()
    torch.manual_seed(5)
    p, _ = O.init_field_params(4, 0.5, 5.0)
    model = _model(p)
    rows = 300
    pts = torch.rand(rows, 3) * 2 - 1
    dirs = torch.nn.functional.normalize(torch.randn(rows, 3), dim=-1)
    x = torch.cat([O.freq_encode(pts, 10), O.freq_encode(dirs, 4)], -1)
    g = torch.randn(rows, 4)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    O.field_mlp(q, x).backward(g)
    out = model(x.to(DEV))
    out.backward(g.to(DEV))
    for name, prm in model.named_parameters():
        st = grad_stats(prm.grad.cpu(), q[name].grad)
        assert st["rel_l2"] <= 0.15 and st["cos"] >= 0.99, (name, st)


def test_mse_loss_grad():
    K = _K()
    torch.manual_seed(0)
    x, t = torch.rand(1000, 3), torch.rand(1000, 3)
    xr = x.clone().requires_grad_(True)
    ref = torch.mean((xr - t) ** 2)
    ref.backward()
    loss, grad = K.mse_loss_grad(x.to(DEV), t.to(DEV))
    assert abs(loss.item() - ref.item()) <= 1e-6
    assert (grad.cpu() - xr.grad).abs().max() <= 1e-9


@pytest.mark.parametrize("name", ["lego_train", "fern_train"])
def test_train_step_gradients_match_reference_fixture(name):
    """render -> loss -> backward through the drop-in surface vs the reference's autograd
    (main.py:376-385) on the same rays, weights, target and random draws.

    delta_last = 1e10 makes a ray's colour a step function of the sign of the far sample's density
    (SURVEY.md App. C); rays whose sign differs between the BF16 and the fp32 forward are excluded
    from BOTH sides (counted and bounded), like the forward parity tests do.  When no ray flips, the
    gradients are also compared with the norms / 96-element heads recorded from the real reference."""
    from cv_nerf_b200 import main as M
    g = golden(f"render_{name}.npz")
    h, w = int(g["hwf"][0]), int(g["hwf"][1])
    f = focal_of(g)
    coarse_p, fine_p = O.init_field_params(int(g["seed"]), float(g["sigma_bias"]), float(g["sigma_gain"]))
    coarse, fine = _model(coarse_p), _model(fine_p)
    noise = float(g["noise"])
    cpu = {k: torch.from_numpy(g[k]) for k in ("u", "t_rand", "rays_o", "rays_d", "target")}
    if noise > 0:
        cpu["noise_c"], cpu["noise_f"] = torch.from_numpy(g["noise_c"]), torch.from_numpy(g["noise_f"])
    kw = dict(white_bkg=bool(g["white_bkg"]), ndc=bool(g["ndc"]), near=float(g["near"]), far=float(g["far"]))

    def ours(keep):
        d = M.RenderDraws(u=cpu["u"][keep], t_rand=cpu["t_rand"][keep])
        if noise > 0:
            d.noise_c, d.noise_f = cpu["noise_c"][keep], cpu["noise_f"][keep]
        rays = torch.stack([cpu["rays_o"][keep], cpu["rays_d"][keep]], 0).to(DEV)
        packed = M.K.pack_rays(h, w, f, rays_o=rays[0], rays_d=rays[1], ndc=kw["ndc"], near=kw["near"], far=kw["far"])
        return M.render_rays(packed, draws=d, extras=True, coarse_model=coarse, fine_model=fine, n_coarse_samples=64,
                             n_fine_samples=128, perturb=1., noise=noise, white_bkg=kw["white_bkg"])

    def oracle(keep, cp, fp):
        d = O.RenderDraws(u=cpu["u"][keep], t_rand=cpu["t_rand"][keep])
        if noise > 0:
            d.noise_c, d.noise_f = cpu["noise_c"][keep] * noise, cpu["noise_f"][keep] * noise
        return O.render_image(h, w, f, cp, fp, rays=(cpu["rays_o"][keep], cpu["rays_d"][keep]), ndc=kw["ndc"],
                              near=kw["near"], far=kw["far"], draws=d, white_bkg=kw["white_bkg"], extras=True)

    everything = torch.ones(cpu["u"].shape[0], dtype=torch.bool)
    with torch.no_grad():
        a, b = ours(everything), oracle(everything, coarse_p, fine_p)
    flip = torch.zeros_like(everything)
    for key in ("raw_c", "raw_f"):
        flip |= (a[key][:, -1, 3].cpu() > 0) != (b[key][:, -1, 3] > 0)
    n_flip = int(flip.sum())
    print(name, "far-sample sign flips:", n_flip, "of", flip.numel())
    assert n_flip <= max(2, flip.numel() // 20)
    keep = ~flip

    target = cpu["target"][keep]
    out = ours(keep)
    loss = torch.mean((out["rgb_map"] - target.to(DEV)) ** 2) + torch.mean((out["rgb_c"] - target.to(DEV)) ** 2)
    loss.backward()
    cq = {k: v.clone().requires_grad_(True) for k, v in coarse_p.items()}
    fq = {k: v.clone().requires_grad_(True) for k, v in fine_p.items()}
    ref = oracle(keep, cq, fq)
    ref_loss = O.train_loss(ref["rgb_map"], ref["rgb_c"], target)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-3 * max(1., ref_loss.item()), (loss.item(), ref_loss.item())

    bad, worst_norm, worst_cos = [], 0., 1.
    all_got = torch.cat([prm.grad.detach().cpu().reshape(-1) for net in (coarse, fine) for _, prm in net.named_parameters()])
    all_ref = torch.cat([q[k].grad.reshape(-1) for net, q in ((coarse, cq), (fine, fq)) for k, _ in net.named_parameters()])
    whole = grad_stats(all_got, all_ref)
    print(name, "whole gradient vector (1 191 688 values) vs the fp32 reference:", whole)
    record("grad_e2e_whole", dict(fixture=name, rel_l2=whole["rel_l2"], cos=whole["cos"], rays=int(keep.sum()), n_flip=n_flip))
    assert whole["rel_l2"] <= 0.04 and whole["cos"] >= 0.999, whole     # measured 0.021 / 0.0039, cos 0.99986 / 0.99999
    for tag, net, q in (("coarse", coarse, cq), ("fine", fine, fq)):
        for k, prm in net.named_parameters():
            st = grad_stats(prm.grad.detach().cpu(), q[k].grad)
            got_n = prm.grad.norm().item()
            rel_norm = abs(got_n - st["ref_norm"]) / max(st["ref_norm"], 1e-30)
            print(tag, k, "norm rel err", rel_norm, st)
            record("grad_e2e", dict(fixture=name, net=tag, tensor=k, norm_rel_err=rel_norm, rel_l2=st["rel_l2"],
                                    cos=st["cos"], n_flip=n_flip))
            worst_norm, worst_cos = max(worst_norm, rel_norm), min(worst_cos, st["cos"])
            # BF16 contractions + ReLU masks of a BF16 forward on <= 96 rays, per tensor: norm within
            # 3 % (measured <= 2.0 %), cos >= 0.985 (measured >= 0.9886), rel-L2 <= 0.18 (measured
            # <= 0.153 at fine l1.weight, the end of the chain); the kernels' own arithmetic is pinned to
            # 2e-3 rel-L2 (measured 6e-4) by test_field_backward_stages
            if rel_norm > 3e-2 or st["cos"] < 0.985 or st["rel_l2"] > 0.18:
                bad.append((tag, k, rel_norm, st))
            if n_flip == 0:
                gn = float(g[f"gnorm/{tag}.{k}"])
                if abs(got_n - gn) > 5e-2 * gn:
                    bad.append((tag, k, "vs recorded reference norm", got_n, gn))
    print("worst gradient-norm error", worst_norm, "worst cosine", worst_cos)
    assert not bad, bad
