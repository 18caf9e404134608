"""Parity on the configurations BASELINE.json quotes the metric on, at their full frame sizes,
through the public ``render(c2w=...)`` call: the GPU renders the whole frame, the CPU oracle (pinned
to the real reference by tests/test_oracle_golden.py) re-renders a pixel subset on the same uniform
draws.  Tolerance (north_star): max-abs rgb <= 1e-2 for BF16 MLP math, far-sample sign-flip rays
 counted separately and bounded; PSNR delta <= 0.1 dB.

(a) lego 800x800 (configs[1], the headline)      main.py:49-87
(b) fern 378x504 with NDC (configs[2])           data_helpers.py:327-344
(c) skull spiral frames 0/30/60/90 (configs[4])  -- frames 0 and 60 have NDC origins ~2.3e8
(d) "sharpened" weights: 300 TrainSteps on a high-frequency target, then (a)'s check on them
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.helpers import flip_aware_stats, golden, load_model_params, psnr, record

pytestmark = pytest.mark.gpu
DEV = "cuda"
RGB_TOL = 1e-2


def _nets(seed, sigma_bias=1.0, sigma_gain=5.0):
    from cv_nerf_b200.model import Model
    cp, fp = O.init_field_params(seed, sigma_bias, sigma_gain)
    return cp, fp, load_model_params(Model(), cp).to(DEV), load_model_params(Model(), fp).to(DEV)


def _frame_vs_oracle(tag, h, w, f, pose, cp, fp, coarse, fine, *, ndc, near, far, white_bkg, n_check, seed,
                     max_flip_frac=0.05, rgb_tol=RGB_TOL, rgb_tol_fine=None, frac_over_1e2=0.0, emulate=False):
    """Full frame on the GPU, `n_check` random pixels of it on the CPU oracle; returns the stats.
    ``emulate``: additionally restate the kernel's BF16 rounding points on the CPU (helpers.emulate_field)
    for the checked rays' coarse samples and require the GPU to match THAT tightly -- separates "this
    is what BF16 tensor-core math gives" from "the kernel is wrong"."""
    from cv_nerf_b200 import main as M
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(h * w, 128, generator=g)
    kw = dict(n_coarse_samples=64, n_fine_samples=128, white_bkg=white_bkg, ndc=ndc, near=near, far=far)
    with torch.no_grad():
        rgb, ex = M.render(h, w, f, c2w=pose.to(DEV), draws=M.RenderDraws(u=u), coarse_model=coarse,
                           fine_model=fine, extras=True, **kw)
    assert rgb.shape == (h, w, 3) and ex["rgb_c"].shape == (h, w, 3)
    assert torch.isfinite(rgb).all() and torch.isfinite(ex["rgb_c"]).all()
    idx = torch.randperm(h * w, generator=g)[:n_check]
    o, d = O.ray_grid(h, w, f, pose)
    ref = O.render_image(h, w, f, cp, fp, rays=(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]),
                         draws=O.RenderDraws(u=u[idx]), n_coarse=64, n_fine=128, white_bkg=white_bkg, ndc=ndc,
                         near=near, far=far, extras=True)
    di = idx.to(DEV)
    # the packed rays of the checked pixels are the reference's, bit for bit
    packed = M.K.pack_rays(h, w, f, pose=pose.to(DEV), ndc=ndc, near=near, far=far)[di].cpu()
    assert torch.equal(packed.view(torch.int32), ref["rays"].view(torch.int32)), "packed rays not bit-exact"
    out = {}
    for key, raw_key in (("rgb_c", "raw_c"), ("rgb_map", "raw_f")):
        got = (rgb if key == "rgb_map" else ex[key]).reshape(-1, 3)[di].cpu()
        sig = ex[raw_key].reshape(h * w, -1, 4)[di, -1, 3].cpu()
        st = flip_aware_stats(got, ref[key], sig, ref[raw_key][:, -1, 3])
        st["psnr_vs_oracle"] = psnr(got, ref[key])
        raw_got = ex[raw_key].reshape(h * w, -1, 4)[di].cpu()
        st["raw_rel_l2"] = ((raw_got - ref[raw_key]).norm() / ref[raw_key].norm().clamp_min(1e-30)).item()
        st["raw_absmax_ref"] = ref[raw_key].abs().max().item()
        print(tag, key, st)
        record("frame_parity", dict(case=tag, output=key, frame=f"{h}x{w}", checked=n_check, **st))
        assert st["max_noflip"] <= (rgb_tol_fine if key == "rgb_map" and rgb_tol_fine else rgb_tol), (tag, key, st)
        assert st["n_gt_1e-2"] - st["n_flip"] <= frac_over_1e2 * n_check, (tag, key, st)
        assert st["n_flip"] <= max(2, int(n_check * max_flip_frac)), (tag, key, st)
        out[key] = (got, ref[key], st)
    if emulate:
        from tests.helpers import emulate_field, vterm_reference
        rays = ref["rays"]
        z_c = ex["z_c"].reshape(h * w, -1)[di].cpu()
        pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z_c[:, :, None]
        s_c = z_c.shape[1]
        vt = vterm_reference(cp, rays[:, 8:11])[:, None].expand(n_check, s_c, 128).reshape(-1, 128)
        raw_emu = emulate_field(cp, O.freq_encode(pts.reshape(-1, 3), 10), vt).reshape(n_check, s_c, 4)
        raw_got = ex["raw_c"].reshape(h * w, s_c, 4)[di].cpu()
        rgb_emu, _ = O.composite(raw_emu, z_c, rays[:, 3:6], None, white_bkg)
        got_c = ex["rgb_c"].reshape(-1, 3)[di].cpu()
        st = {"raw_max_abs_vs_emulation": (raw_got - raw_emu).abs().max().item(),
              "raw_rel_l2_vs_emulation": ((raw_got - raw_emu).norm() / raw_emu.norm()).item(),
              "rgb_c_max_abs_vs_emulation": (got_c - rgb_emu).abs().max().item(),
              "rgb_c_emulation_vs_fp32_max_abs": (rgb_emu - ref["rgb_c"]).abs().max().item()}
        # fine pass on the GPU's own fine depths (the resampling amplifies the coarse differences: a sample
        # lands on the other side of a density edge), so that only the field's arithmetic is compared
        z_f = ex["z_f"].reshape(h * w, -1)[di].cpu()
        s_f = z_f.shape[1]
        pts_f = rays[:, None, 0:3] + rays[:, None, 3:6] * z_f[:, :, None]
        vt_f = vterm_reference(fp, rays[:, 8:11])[:, None].expand(n_check, s_f, 128).reshape(-1, 128)
        raw_f_emu = emulate_field(fp, O.freq_encode(pts_f.reshape(-1, 3), 10), vt_f).reshape(n_check, s_f, 4)
        raw_f_fp32 = O.query_field(fp, pts_f, rays[:, 8:11])
        rgb_f_emu, _ = O.composite(raw_f_emu, z_f, rays[:, 3:6], None, white_bkg)
        rgb_f_fp32, _ = O.composite(raw_f_fp32, z_f, rays[:, 3:6], None, white_bkg)
        got_f = rgb.reshape(-1, 3)[di].cpu()
        st.update({"rgb_f_max_abs_vs_emulation_same_depths": (got_f - rgb_f_emu).abs().max().item(),
                   "rgb_f_max_abs_vs_fp32_same_depths": (got_f - rgb_f_fp32).abs().max().item(),
                   "rgb_f_frac_gt_1e-2_vs_fp32_same_depths": ((got_f - rgb_f_fp32).abs().amax(-1) > 1e-2).float().mean().item()})
        print(tag, "vs the CPU emulation of BF16 tensor-core math:", st)
        record("frame_vs_bf16_emulation", dict(case=tag, **st))
        out["emulation"] = st
    # north_star's PSNR criterion against a synthetic target correlated with the image
    got, want, _ = out["rgb_map"]
    target = torch.rand(n_check, 3, generator=g) * 0.3 + 0.35 * want + 0.2
    delta = abs(psnr(got, target) - psnr(want, target))
    print(tag, "PSNR delta vs a synthetic target", delta, "dB")
    record("frame_psnr_delta", dict(case=tag, delta_db=delta))
    assert delta <= 0.1
    return out


def test_lego_800x800_full_frame_parity():
    """The headline configuration (BASELINE.json configs[1]): 640 000 rays through render(c2w=...)."""
    cp, fp, coarse, fine = _nets(0)
    pose = O.lego_pose(36., -30., 4.)[:3, :4]
    _frame_vs_oracle("lego800", 800, 800, 1111.1110311937682, pose, cp, fp, coarse, fine, ndc=False, near=2.,
                     far=6., white_bkg=True, n_check=2048, seed=11)


def test_fern_378x504_full_frame_ndc_parity():
    """configs[2]: fern intrinsics (np.float32 focal) with NDC rays.  The fern poses are not in the
    reference tree; the pose is a recentred training pose of the skull capture."""
    cp, fp, coarse, fine = _nets(2)
    sk = golden("skull_spiral.npz")
    pose = torch.from_numpy(sk["train_poses"][3][:3, :4]).float()
    _frame_vs_oracle("fern", 378, 504, np.float32(407.5657), pose, cp, fp, coarse, fine, ndc=True, near=0.,
                     far=1., white_bkg=False, n_check=2048, seed=12)


@pytest.mark.parametrize("frame", [0, 30, 60, 90])
def test_skull_spiral_frames_parity(frame):
    """configs[4]: the reference's own spiral poses.  Frames 0 and 60 put the NDC origin at ~2.3e8
    (every |raw| is ~1e6 and every colour saturates); the BF16 field must still return the
    reference's pixels -- measured here, not assumed."""
    cp, fp, coarse, fine = _nets(2)
    sk = golden("skull_spiral.npz")
    h, w, f = int(sk["hwf"][0]), int(sk["hwf"][1]), np.float32(sk["hwf"][2])
    pose = torch.from_numpy(sk["render_poses"][frame]).float()
    _frame_vs_oracle(f"skull_f{frame:03d}", h, w, f, pose, cp, fp, coarse, fine, ndc=True, near=0., far=1.,
                     white_bkg=False, n_check=1024, seed=100 + frame)


def test_sharpened_weights_parity():
    """Parity is not a property of smooth random-init weights only: fit both networks for 300 steps
    to a target with sharp edges (a 0/1 checkerboard of 40-pixel squares with a different colour
    phase per channel, one view), then render that 400x400 view with the trained weights and check a
    pixel subset against the oracle running the SAME trained weights in fp32."""
    from cv_nerf_b200.train import TrainStep
    torch.manual_seed(7)
    h = w = 400
    f = 555.5555155968841
    cp, fp, coarse, fine = _nets(5, 0.5, 10.0)
    ii, jj = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    image = torch.stack([(((ii + s) // 40 + (jj + 2 * s) // 40) % 2).float() for s in (0, 13, 26)], -1).to(DEV)
    pose = O.lego_pose(-180., -30., 4.)[:3, :4]
    ts = TrainStep(coarse, fine, height=h, width=w, focal=f, n_rays=4096, perturb=1., noise=0., white_bkg=True,
                   ndc=False, near=2., far=6., lr=1e-3, lr_decay=250, seed=3)
    pose_dev = pose.to(DEV)
    losses = [ts.step(image, pose_dev).clone() for _ in range(300)]      # step() returns its own loss buffer
    first, last = torch.stack(losses[:5]).mean().item(), torch.stack(losses[-5:]).mean().item()
    print("sharpening: loss", first, "->", last)
    assert last < 0.6 * first, "the fit did not sharpen the field"
    trained = []
    for net, p0 in ((coarse, cp), (fine, fp)):
        sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        assert max((sd[k] - p0[k]).abs().max().item() for k in p0) > 1e-2, "training did not move the weights"
        trained.append(sd)
    wmax = max(v.abs().max().item() for sd in trained for v in sd.values())
    # A sharp field (|sigma| up to ~130 here) amplifies the BF16 rounding of the activations: SURVEY.md
    # App. C measured max-abs 3.6e-2 (excluding flips) for a BF16 EMULATION OF THE REFERENCE on a sharp
    # density head, i.e. the north_star's 1e-2 is not attainable by BF16 MLP math on such weights.  So:
    # (1) against the fp32 oracle: <= 1e-2 for at least 95 % of the non-flip rays; every coarse pixel within
    #     "what the BF16 format costs on these weights (CPU emulation vs fp32) + 1e-2" (measured 2.2e-2..4.5e-2
    #     over the training trajectories of this round, the emulation's own distance to fp32 within 1e-4 of
    #     it); the fine pass additionally moves its samples (inverse-CDF
    #     resampling of slightly different coarse weights puts a sample on the other side of a density
    #     edge): measured 0.12..0.22 on the worst of 2048 pixels over three training runs (the training
    #     accumulates with floating-point atomics, so the trained weights differ from run to run), bounded
    #     at 0.35; on the SAME fine depths the fine pass is within 2.3e-2 of fp32, 0.7 % of the rays > 1e-2;
    # (2) PSNR delta <= 0.1 dB (north_star's second criterion; asserted inside);
    # (3) the GPU agrees with the CPU emulation of BF16 tensor-core math to 1e-2 in the two passes (the fine
    #     pass on the GPU's own depths) -- what is left against fp32 is the number format, not the kernel.
    out = _frame_vs_oracle("lego400_trained", h, w, f, pose, trained[0], trained[1], coarse, fine, ndc=False, near=2.,
                           far=6., white_bkg=True, n_check=2048, seed=13, rgb_tol=0.1, rgb_tol_fine=0.35,
                           frac_over_1e2=0.05, emulate=True)
    emu = out["emulation"]
    assert out["rgb_c"][2]["max_noflip"] <= emu["rgb_c_emulation_vs_fp32_max_abs"] + 1e-2, (out["rgb_c"][2], emu)
    # (the CPU emulation takes torch's sin/cos; the kernel's own encoding can land on the neighbouring BF16
    # value, which a sharp field amplifies: measured 1e-3 .. 6e-3 over four training runs)
    assert emu["rgb_c_max_abs_vs_emulation"] <= 1e-2, emu
    assert emu["rgb_f_max_abs_vs_emulation_same_depths"] <= 1e-2, emu
    assert emu["raw_rel_l2_vs_emulation"] <= 1e-3, emu                         # measured 9.8e-5
    assert emu["rgb_f_max_abs_vs_fp32_same_depths"] <= 0.35, emu               # measured 2.3e-2 .. 0.11 (worst ray)
    assert emu["rgb_f_frac_gt_1e-2_vs_fp32_same_depths"] <= 0.05, emu
    got, want, _ = out["rgb_map"]
    spread = want.std().item()
    record("sharpened_weights", dict(loss_first=first, loss_last=last, steps=300, weight_absmax=wmax,
                                     rendered_rgb_std=spread))
    assert spread > 0.1, "the trained scene renders flat: the parity check would be vacuous"
