"""End-to-end parity of the drop-in render path (cv_nerf_b200.main.render) against fixtures
recorded from the real reference and against the CPU oracle, on identical rays, weights and
random draws.  Tolerance (BASELINE.json north_star): max-abs rgb error <= 1e-2 for BF16 MLP math,
with far-sample sign-flip rays counted separately."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.helpers import flip_aware_stats, focal_of, golden, load_model_params, psnr, record

pytestmark = pytest.mark.gpu
DEV = "cuda"
RGB_TOL = 1e-2


def _models(g):
    from cv_nerf_b200.model import Model
    sb = None if np.isnan(g["sigma_bias"]) else float(g["sigma_bias"])
    coarse_p, fine_p = O.init_field_params(int(g["seed"]), sb, float(g["sigma_gain"]))
    coarse = load_model_params(Model(), coarse_p).to(DEV)
    fine = load_model_params(Model(), fine_p).to(DEV)
    return coarse_p, fine_p, coarse, fine


def _draws(g):
    from cv_nerf_b200.main import RenderDraws
    d = RenderDraws(u=torch.from_numpy(g["u"]))
    od = O.RenderDraws(u=torch.from_numpy(g["u"]))
    noise = float(g["noise"])
    if bool(g["train"]):
        d.t_rand = od.t_rand = torch.from_numpy(g["t_rand"])
        if noise > 0:
            d.noise_c, d.noise_f = torch.from_numpy(g["noise_c"]), torch.from_numpy(g["noise_f"])
            od.noise_c, od.noise_f = d.noise_c * noise, d.noise_f * noise
    return d, od


def _render_both(name):
    from cv_nerf_b200 import main as M
    g = golden(f"render_{name}.npz")
    h, w = int(g["hwf"][0]), int(g["hwf"][1])
    f = focal_of(g)
    coarse_p, fine_p, coarse, fine = _models(g)
    draws, odraws = _draws(g)
    train = bool(g["train"])
    kw = dict(coarse_model=coarse, fine_model=fine, q_fn=None, n_coarse_samples=64, n_fine_samples=128,
              perturb=1. if train else 0., noise=float(g["noise"]), white_bkg=bool(g["white_bkg"]),
              ndc=bool(g["ndc"]), near=float(g["near"]), far=float(g["far"]))
    rays = torch.stack([torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])], 0).to(DEV)
    return g, kw, rays, draws, (coarse_p, fine_p, odraws, h, w, f)


# skull frames 0 and 60: NDC origins of ~2.3e8 -> |raw| ~ 1e6, every colour saturates
DEGENERATE = ("lego_test_stock", "skull_f000_test", "skull_f060_test")


@pytest.mark.parametrize("name", ["lego_test", "fern_test", "lego_train", "fern_train", "lego_test_stock",
                                  "lego800_test", "skull_f000_test", "skull_f030_test", "skull_f060_test",
                                  "skull_f090_test"])
def test_render_matches_reference_fixture(name):
    from cv_nerf_b200 import main as M
    g, kw, rays, draws, (coarse_p, fine_p, odraws, h, w, f) = _render_both(name)
    with torch.no_grad():
        rgb, extras = M.render(h, w, f, rays=rays, draws=draws, **kw)
    ref = O.render_image(h, w, f, coarse_p, fine_p, rays=(torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])),
                         ndc=kw["ndc"], near=kw["near"], far=kw["far"], draws=odraws, white_bkg=kw["white_bkg"],
                         extras=True)
    with torch.no_grad():
        packed = M.K.pack_rays(h, w, f, rays_o=rays[0], rays_d=rays[1], ndc=kw["ndc"], near=kw["near"], far=kw["far"])
        ours = M.render_rays(packed, draws=draws, extras=True, **{k: v for k, v in kw.items()
                                                                 if k not in ("ndc", "near", "far")})
    assert torch.equal(packed.cpu().view(torch.int32), ref["rays"].view(torch.int32)), "packed rays not bit-exact"
    assert torch.equal(ours["z_c"].cpu().view(torch.int32), ref["z_c"].contiguous().view(torch.int32))
    for key, want, sig_key in (("rgb_c", g["rgb_c"], "raw_c"), ("rgb_map", g["rgb_map"], "raw_f")):
        got = (extras[key] if key == "rgb_c" else rgb).cpu()
        st = flip_aware_stats(got, torch.from_numpy(want), ours[sig_key][:, -1, 3].cpu(), ref[sig_key][:, -1, 3])
        print(name, key, st, "psnr_vs_ref", psnr(got, torch.from_numpy(want)))
        record("render_fixture", dict(fixture=name, output=key, rays=int(got.shape[0]), **st))
        assert st["max_noflip"] <= RGB_TOL, (name, key, st)
        assert st["n_flip"] <= max(2, got.shape[0] // 20), (name, key, st)
    if name not in DEGENERATE:
        # opacity accumulated before the far sample (delta_last = 1e10 makes that one absorb the rest)
        acc = ref["w_f"][:, :-1].sum(-1).mean().item()
        assert 0.05 < acc < 0.999, f"test scene is degenerate (acc {acc})"


def test_render_c2w_full_image_and_row_shards():
    """render(c2w=...) on a small image equals the oracle, and rendering row blocks separately
    (how a frame is sharded over GPUs) reproduces the same pixels bit for bit."""
    from cv_nerf_b200 import main as M
    from cv_nerf_b200.model import Model
    h, w, f = 24, 32, 40.0
    coarse_p, fine_p = O.init_field_params(3, 1.0, 5.0)
    coarse = load_model_params(Model(), coarse_p).to(DEV)
    fine = load_model_params(Model(), fine_p).to(DEV)
    pose = O.lego_pose(20., -30., 4.)[:3, :4]
    u = torch.rand(h * w, 128, generator=torch.Generator().manual_seed(1))
    kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=True,
              ndc=False, near=2., far=6.)
    with torch.no_grad():
        rgb, extras = M.render(h, w, f, c2w=pose.to(DEV), draws=M.RenderDraws(u=u), **kw)
        parts = []
        for r0, r1 in ((0, 7), (7, 16), (16, 24)):
            part, _ = M.render(h, w, f, c2w=pose.to(DEV), rows=(r0, r1),
                               draws=M.RenderDraws(u=u[r0 * w:r1 * w]), **kw)
            parts.append(part)
    assert rgb.shape == (h, w, 3) and extras["rgb_c"].shape == (h, w, 3)
    assert torch.equal(torch.cat(parts, 0), rgb), "row-sharded render differs from the full render"
    ref = O.render_image(h, w, f, coarse_p, fine_p, c2w=pose, ndc=False, near=2., far=6.,
                         draws=O.RenderDraws(u=u), white_bkg=True)
    assert (rgb.cpu() - ref["rgb_map"]).abs().max().item() <= RGB_TOL
    assert (extras["rgb_c"].cpu() - ref["rgb_c"]).abs().max().item() <= RGB_TOL


def test_drop_in_surface_small_calls():
    """The reference-named helpers work stand-alone on CUDA tensors and refuse CPU tensors."""
    import cv_nerf_b200
    from cv_nerf_b200 import main as M, model as MD, utils as U, data_helpers as DH
    pose = O.lego_pose()[:3, :4]
    o, d = M.compute_rays(16, 20, 30.0, pose.to(DEV))
    o_ref, d_ref = O.ray_grid(16, 20, 30.0, pose)
    assert torch.equal(d.cpu(), d_ref) and torch.equal(o.cpu(), o_ref.contiguous())
    on, dn = DH.get_ndc(16, 20, 30.0, 1., o, d)
    on_ref, dn_ref = O.ndc_warp(16, 20, 30.0, 1., o_ref, d_ref)
    assert torch.equal(on.cpu(), on_ref) and torch.equal(dn.cpu(), dn_ref)
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        M.compute_rays(16, 20, 30.0, pose)          # CPU tensor: no fallback
    g = golden("units.npz")
    s = U.inv_transform_sampling(torch.from_numpy(g["pdf_bins"]).to(DEV), torch.from_numpy(g["pdf_w"]).to(DEV), 128,
                                 u=torch.from_numpy(g["pdf_u"]).to(DEV))
    np.testing.assert_allclose(s.cpu().numpy(), g["pdf_samples"], rtol=0, atol=2e-5)
    raw, z, dd = (torch.from_numpy(g[f"comp_c_{k}"]).to(DEV) for k in ("raw", "z", "d"))
    rgb, wts = M.process_volume_info(raw, z, dd, 0., True)
    np.testing.assert_allclose(rgb.cpu().numpy(), g["comp_c_rgb_white"], rtol=0, atol=2e-6)
    # Model.forward on an embedded [.,90] batch and net_forward on points agree with the oracle
    coarse_p, _ = O.init_field_params(0, 1.0, 5.0)
    net = load_model_params(MD.Model(), coarse_p).to(DEV)
    pts = torch.randn(5, 64, 3, generator=torch.Generator().manual_seed(8))
    dirs = torch.nn.functional.normalize(torch.randn(5, 3, generator=torch.Generator().manual_seed(9)), dim=-1)
    want = O.query_field(coarse_p, pts, dirs)
    with torch.no_grad():
        got = MD.net_forward(pts.to(DEV), dirs.to(DEV), net, MD.FreqEmbedding(10).embed, MD.FreqEmbedding(4).embed)
        x = torch.cat([MD.FreqEmbedding(10).embed(pts.to(DEV).reshape(-1, 3)),
                       MD.FreqEmbedding(4).embed(dirs.to(DEV)[:, None].expand(5, 64, 3).reshape(-1, 3))], -1)
        got2 = net(x).reshape(5, 64, 4)
    assert (got.cpu() - want).abs().max().item() <= 3e-2
    assert (got2.cpu() - want).abs().max().item() <= 3e-2


def test_depth_acc_disp_maps_and_to_byte():
    """Extra maps (depth, acc, disp: nerf-pytorch raw2outputs definitions over the reference's
    weights) and the device-side to_byte against numpy."""
    from cv_nerf_b200 import kernels as K
    torch.manual_seed(1)
    n, s = 300, 192
    w = torch.rand(n, s) / s
    z = torch.sort(torch.rand(n, s) * 4 + 2, -1).values
    m = K.composite_maps(w.to(DEV), z.to(DEV)).cpu()
    depth, acc = (w * z).sum(-1), w.sum(-1)
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    assert (m[:, 0] - depth).abs().max() <= 1e-5 and (m[:, 1] - acc).abs().max() <= 1e-6
    assert ((m[:, 2] - disp).abs() / disp).max() <= 1e-5
    x = torch.cat([torch.rand(1001, 3) * 1.4 - 0.2, torch.tensor([[0., 1., 0.5]])])
    got = K.to_byte(x.to(DEV)).cpu().numpy()
    want = (255 * np.clip(x.numpy(), 0, 1)).astype(np.uint8)
    assert np.array_equal(got, want)


def test_render_full_video_pipeline():
    """render_full (main.py:102-124): float and uint8 stacks agree with per-frame render()."""
    from cv_nerf_b200 import main as M
    from cv_nerf_b200.model import Model
    coarse_p, fine_p = O.init_field_params(0, 1.0, 5.0)
    coarse, fine = load_model_params(Model(), coarse_p).to(DEV), load_model_params(Model(), fine_p).to(DEV)
    kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=True,
              ndc=False, near=2., far=6., perturb=False, noise=0.)
    poses = [O.lego_pose(a, -30., 4.).to(DEV) for a in (-180., -90., 0.)]
    torch.manual_seed(5)
    vid = M.render_full(poses, [20, 24, 30.], 32768, kw, verbose=False)
    torch.manual_seed(5)
    vid8 = M.render_full(poses, [20, 24, 30.], 32768, kw, as_bytes=True, verbose=False)
    assert vid.shape == (3, 20, 24, 3) and vid8.dtype == np.uint8
    assert np.array_equal(vid8, (255 * np.clip(vid, 0, 1)).astype(np.uint8))
    torch.manual_seed(5)
    with torch.no_grad():
        first, _ = M.render(20, 24, 30., c2w=poses[0][:3, :4], **kw)
    assert np.array_equal(first.cpu().numpy(), vid[0])
    half = M.render_full(poses[:1], [20, 24, 30.], 32768, kw, factor=2, verbose=False)
    assert half.shape == (1, 10, 12, 3)


def test_full_frame_lego_400_parity_on_pixel_subset():
    """BASELINE.json configs[0] at full size (lego half-res 400x400, 64+128 samples, white background):
    the whole frame is rendered through render(c2w=...); the CPU oracle, which needs minutes for the
    full frame, checks 2048 of its pixels on the same uniform draws.  Also the north_star's PSNR
    criterion: |PSNR(ours, target) - PSNR(reference, target)| <= 0.1 dB against a synthetic target."""
    from cv_nerf_b200 import main as M
    from cv_nerf_b200.model import Model
    h = w = 400
    f = 555.5555155968841
    coarse_p, fine_p = O.init_field_params(0, 1.0, 5.0)
    coarse, fine = load_model_params(Model(), coarse_p).to(DEV), load_model_params(Model(), fine_p).to(DEV)
    pose = O.lego_pose(-180., -30., 4.)[:3, :4]
    g = torch.Generator().manual_seed(11)
    u = torch.rand(h * w, 128, generator=g)
    kw = dict(n_coarse_samples=64, n_fine_samples=128, white_bkg=True, ndc=False, near=2., far=6.)
    with torch.no_grad():
        rgb, extras = M.render(h, w, f, c2w=pose.to(DEV), draws=M.RenderDraws(u=u), coarse_model=coarse,
                               fine_model=fine, **kw)
    assert rgb.shape == (h, w, 3)
    idx = torch.randperm(h * w, generator=g)[:2048]
    o, d = O.ray_grid(h, w, f, pose)
    ref = O.render_image(h, w, f, coarse_p, fine_p, rays=(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]),
                         draws=O.RenderDraws(u=u[idx]), n_coarse=64, n_fine=128, white_bkg=True, ndc=False,
                         near=2., far=6.)
    got = rgb.reshape(-1, 3)[idx.to(DEV)].cpu()
    got_c = extras["rgb_c"].reshape(-1, 3)[idx.to(DEV)].cpu()
    err, err_c = (got - ref["rgb_map"]).abs().max().item(), (got_c - ref["rgb_c"]).abs().max().item()
    print("400x400 subset: max|rgb - ref| fine", err, "coarse", err_c, "PSNR(ours vs ref)", psnr(got, ref["rgb_map"]))
    assert err <= RGB_TOL and err_c <= RGB_TOL
    target = torch.rand(2048, 3, generator=g) * 0.3 + 0.35 * ref["rgb_map"] + 0.2
    delta = abs(psnr(got, target) - psnr(ref["rgb_map"], target))
    print("PSNR delta vs a synthetic target", delta, "dB")
    assert delta <= 0.1
