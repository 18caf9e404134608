"""GPU tests of the train iteration (main.py:344-394): fused Adam against torch.optim.Adam, the
device-side ray/target batch against a gather from the full ray grid, TrainStep against the autograd
path, and a short optimisation run."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.helpers import focal_of, golden, grad_stats, load_model_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _K():
    import cv_nerf_b200
    return cv_nerf_b200.kernels


def test_fused_adam_matches_torch_adam():
    from cv_nerf_b200.train import FusedAdam
    torch.manual_seed(0)
    shapes = [(256, 63), (256,), (1, 256), (1,), (3, 128), (128, 283)]
    ref_p = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=5e-4, betas=(0.9, 0.999))
    ours = FusedAdam(our_p, lr=5e-4, betas=(0.9, 0.999))
    for step in range(5):
        for a, b in zip(ref_p, our_p):
            g = torch.randn(a.shape) * (10. ** (step - 2))
            a.grad, b.grad = g, g.to(DEV)
        ref.step(); ours.step()
        for group in ref.param_groups:           # main.py:392-394 style schedule
            group['lr'] = 5e-4 * 0.1 ** ((step + 1) / 250000)
        for group in ours.param_groups:
            group['lr'] = 5e-4 * 0.1 ** ((step + 1) / 250000)
    for a, b in zip(ref_p, our_p):
        # same formula, different fusion of the fp32 roundings
        assert (a.detach() - b.detach().cpu()).abs().max().item() <= 2e-6 * max(1., a.abs().max().item())


def test_train_rays_equal_gather_from_full_grid():
    K = _K()
    g = golden("rays.npz")
    for case, ndc, near, far in (("lego400", False, 2., 6.), ("fern", True, 0., 1.)):
        h, w = int(g[f"{case}_hwf"][0]), int(g[f"{case}_hwf"][1])
        f = focal_of(g, case + "_")
        pose = torch.from_numpy(g[f"{case}_pose"]).to(DEV)
        full = K.pack_rays(h, w, f, pose=pose, ndc=ndc, near=near, far=far)
        image = torch.rand(h, w, 3, device=DEV)
        pix = torch.randint(0, h * w, (777,), device=DEV, dtype=torch.int32)
        rays, target, _ = K.train_rays(h, w, f, pose, 0, pix=pix, image=image, ndc=ndc, near=near, far=far)
        assert torch.equal(rays.view(torch.int32), full[pix.long()].view(torch.int32)), "batch rays differ from the grid"
        assert torch.equal(target, image.view(-1, 3)[pix.long()])
        # device-side draw: distinct pixels inside the crop window, different per seed
        crop = (h // 4, w // 4, h // 2, w // 2)
        n = 4096
        rays2, tgt2, pix2 = K.train_rays(h, w, f, pose, n, seed=7, crop=crop, image=image, ndc=ndc, near=near,
                                         far=far, want_pix=True)
        p = pix2.long().cpu()
        assert p.unique().numel() == n, "sampling must be without replacement"
        i, j = p // w, p % w
        assert i.min() >= crop[0] and i.max() < crop[0] + crop[2] and j.min() >= crop[1] and j.max() < crop[1] + crop[3]
        assert torch.equal(rays2.view(torch.int32), full[pix2.long()].view(torch.int32))
        _, _, pix3 = K.train_rays(h, w, f, pose, n, seed=8, crop=crop, want_pix=True, ndc=ndc, near=near, far=far)
        assert not torch.equal(pix2, pix3)
        # the whole image can be drawn (a permutation)
        _, _, pall = K.train_rays(16, 24, f, pose, 16 * 24, seed=1, want_pix=True)
        assert sorted(pall.cpu().tolist()) == list(range(16 * 24))


@pytest.mark.parametrize("name", ["lego_train", "fern_train"])
def test_train_step_equals_autograd_path_and_adam(name):
    """TrainStep (no autograd) produces the gradients of the autograd path (same kernels) and the
    parameter update of torch.optim.Adam on those gradients."""
    from cv_nerf_b200 import main as M
    from cv_nerf_b200.model import Model
    from cv_nerf_b200.train import TrainStep
    K = _K()
    g = golden(f"render_{name}.npz")
    h, w = int(g["hwf"][0]), int(g["hwf"][1])
    f = focal_of(g)
    coarse_p, fine_p = O.init_field_params(int(g["seed"]), float(g["sigma_bias"]), float(g["sigma_gain"]))
    mk = lambda p: load_model_params(Model(), p).to(DEV)
    coarse, fine = mk(coarse_p), mk(fine_p)
    coarse2, fine2 = mk(coarse_p), mk(fine_p)
    noise = float(g["noise"])
    draws = M.RenderDraws(u=torch.from_numpy(g["u"]), t_rand=torch.from_numpy(g["t_rand"]))
    if noise > 0:
        draws.noise_c, draws.noise_f = torch.from_numpy(g["noise_c"]), torch.from_numpy(g["noise_f"])
    ro, rd = torch.from_numpy(g["rays_o"]).to(DEV), torch.from_numpy(g["rays_d"]).to(DEV)
    target = torch.from_numpy(g["target"]).to(DEV)
    n = ro.shape[0]
    kw = dict(white_bkg=bool(g["white_bkg"]), ndc=bool(g["ndc"]), near=float(g["near"]), far=float(g["far"]))

    # autograd path
    rgb, extras = M.render(h, w, f, rays=torch.stack([ro, rd], 0), draws=draws, coarse_model=coarse, fine_model=fine,
                           q_fn=None, n_coarse_samples=64, n_fine_samples=128, perturb=1., noise=noise, **kw)
    loss = torch.mean((rgb - target) ** 2) + torch.mean((extras["rgb_c"] - target) ** 2)
    loss.backward()

    # TrainStep path on copies of the networks
    ts = TrainStep(coarse2, fine2, height=h, width=w, focal=f, n_rays=n, perturb=1., noise=noise, lr=5e-4, **kw)
    rays = K.pack_rays(h, w, f, rays_o=ro, rays_d=rd, ndc=kw["ndc"], near=kw["near"], far=kw["far"])
    loss2 = ts.forward_backward(rays, target, draws)
    assert abs(loss2.item() - loss.item()) <= 1e-6 * max(1., abs(loss.item()))
    for idx, net in enumerate((coarse, fine)):
        for got, prm in zip(ts.gradients(idx), net.ordered_params()):
            st = grad_stats(got.cpu(), prm.grad.cpu())
            assert st["rel_l2"] <= 1e-5, st          # identical kernels; only atomics order differs

    # one optimizer step: torch.optim.Adam on the autograd gradients vs the blob Adam
    ref_opt = torch.optim.Adam(list(coarse.parameters()) + list(fine.parameters()), lr=5e-4, betas=(0.9, 0.999))
    ref_opt.step()
    ts.apply_gradients()
    for a, b in ((coarse, coarse2), (fine, fine2)):
        for pa, pb in zip(a.ordered_params(), b.ordered_params()):
            # Adam's first step moves every element by ~lr regardless of the gradient's size, so
            # elements whose tiny gradients differ in sign legitimately differ by 2 lr
            d = (pa.detach() - pb.detach()).abs()
            assert d.max().item() <= 2.1 * 5e-4
            assert (d > 1e-5).float().mean().item() <= 0.02
    assert abs(ts.lr - 5e-4 * 0.1 ** (1 / 250000)) < 1e-12
    # the packed weights follow the update: apply_gradients re-packed both networks in one launch,
    # and the blobs equal a fresh stand-alone pack of the updated parameters
    for net in (coarse2, fine2):
        assert net._packed_key == net._cache_key()
        assert torch.equal(net.packed(), K.pack_model(net.ordered_params()))
        assert torch.equal(net.packed_bwd(), K.pack_model_bwd(net.ordered_params()))


def test_short_training_run_reduces_loss():
    """40 iterations on a synthetic 64x64 target with the device-side batch: the loss goes down."""
    from cv_nerf_b200.model import Model
    from cv_nerf_b200.train import TrainStep
    torch.manual_seed(0)
    coarse_p, fine_p = O.init_field_params(0, 0.5, 5.0)
    coarse, fine = load_model_params(Model(), coarse_p).to(DEV), load_model_params(Model(), fine_p).to(DEV)
    h = w = 64
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    image = torch.stack([xx, yy, 0.5 * torch.ones_like(xx)], -1).to(DEV)
    pose = O.lego_pose(-180., -30., 4.)[:3, :4].to(DEV)
    ts = TrainStep(coarse, fine, height=h, width=w, focal=90., n_rays=1024, perturb=1., noise=0., white_bkg=True,
                   ndc=False, near=2., far=6., lr=5e-4, seed=3)
    losses = [ts.step(image, pose).item() for _ in range(40)]
    print("loss first/last", losses[0], losses[-1])
    assert all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5])


def test_reference_train_loop_body_runs_unchanged():
    """The body of the reference's train loop (main.py:344-394), line for line, on the drop-in
    modules: create_model kwargs, render(rays=...), mse on rgb and rgb_c, backward, optimizer.step,
    learning-rate decay through param_groups."""
    from types import SimpleNamespace
    from cv_nerf_b200 import main as M
    torch.manual_seed(0)
    np.random.seed(0)
    args = SimpleNamespace(netchunk=65536, lr=5e-4, perturb=1., n_fine_samples=128, n_coarse_samples=64,
                           white_bkg=True, noise=0., dtype='blender', no_ndc=False, lr_decay=500, n_rays=512)
    render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer = M.create_model(args)
    assert len(grad_vars) == 48 and render_kwargs_train['ndc'] is False
    render_kwargs_train.update(near=2., far=6.)
    height, width, focal = 40, 40, 55.
    yy, xx = torch.meshgrid(torch.linspace(0, 1, height), torch.linspace(0, 1, width), indexing="ij")
    im = torch.stack([xx, yy, 0.5 * torch.ones_like(xx)], -1).to(DEV)
    pose = O.lego_pose(-180., -30., 4.).to(DEV)
    mse = lambda x, y: torch.mean((x - y) ** 2)
    losses = []
    for i in range(1, 13):
        ray_ori, ray_dir = M.compute_rays(height, width, focal, pose[:3, :4])                     # main.py:351
        ys, xs = torch.meshgrid(torch.linspace(0, height - 1, height), torch.linspace(0, width - 1, width), indexing="ij")
        img_grid = torch.reshape(torch.stack([ys, xs], -1), [-1, 2])                              # main.py:364-367
        idxs = np.random.choice(img_grid.shape[0], size=[args.n_rays], replace=False)             # main.py:368
        selected_pixels = img_grid[idxs].long().to(DEV)
        ray_ori = ray_ori[selected_pixels[:, 0], selected_pixels[:, 1]]                           # main.py:371-372
        ray_dir = ray_dir[selected_pixels[:, 0], selected_pixels[:, 1]]
        batch_rays = torch.stack([ray_ori, ray_dir], 0)
        pixels = im[selected_pixels[:, 0], selected_pixels[:, 1]]
        rgb, extras = M.render(height, width, focal, 32768, rays=batch_rays, **render_kwargs_train)   # main.py:376
        optimizer.zero_grad()                                                                     # main.py:379
        loss = mse(rgb, pixels)
        if 'rgb_c' in extras:
            loss = loss + mse(extras['rgb_c'], pixels)
        loss.backward()
        optimizer.step()
        new_lrate = M.decayed_learning_rate(i, args.lr_decay * 1000, args.lr)                     # main.py:392-394
        for param_group in optimizer.param_groups:
            param_group['lr'] = new_lrate
        losses.append(loss.item())
    assert all(np.isfinite(losses))
    assert np.mean(losses[-3:]) < np.mean(losses[:3])
    # test-time render with the test kwargs (perturb False, noise 0) on a full small frame
    render_kwargs_test.update(near=2., far=6.)
    with torch.no_grad():
        frame, _ = M.render(height, width, focal, 32768, c2w=pose[:3, :4], **render_kwargs_test)
    assert frame.shape == (height, width, 3) and bool(torch.isfinite(frame).all())


def test_deterministic_accumulation_is_bitwise_reproducible():
    """kernels.set_deterministic(True): ordered per-CTA partial sums instead of floating-point atomics.  Two
    identical runs of three train steps give bit-identical parameters (the default mode does not), and the
    deterministic gradients equal the atomic ones to accumulation-order noise."""
    from cv_nerf_b200 import kernels as K
    from cv_nerf_b200.model import Model
    from cv_nerf_b200.train import TrainStep
    h = w = 64
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    image = torch.stack([xx, yy, 0.5 * torch.ones_like(xx)], -1).to(DEV)
    pose = O.lego_pose(-180., -30., 4.)[:3, :4].to(DEV)

    def run(det, steps=3):
        old = K.set_deterministic(det)
        try:
            coarse_p, fine_p = O.init_field_params(0, 0.5, 5.0)
            coarse, fine = load_model_params(Model(), coarse_p).to(DEV), load_model_params(Model(), fine_p).to(DEV)
            ts = TrainStep(coarse, fine, height=h, width=w, focal=90., n_rays=2048, perturb=1., noise=0.5, white_bkg=True,
                           ndc=False, near=2., far=6., lr=5e-4, seed=3)
            losses = [ts.step(image, pose).item() for _ in range(steps)]
            blob = ts.blob.clone()
            params = torch.cat([p.detach().reshape(-1) for p in list(coarse.parameters()) + list(fine.parameters())])
            return losses, blob, params
        finally:
            K.set_deterministic(old)
    l1, b1, p1 = run(True)
    l2, b2, p2 = run(True)
    assert l1 == l2 and torch.equal(b1, b2) and torch.equal(p1, p2), "deterministic mode is not reproducible"
    l3, b3, p3 = run(False, steps=1)
    l4, b4, p4 = run(True, steps=1)
    g = K.grad_blob_floats() - 128 * 256          # the parameter gradients (the scratch region behind them holds G)
    rel = ((b3[:, :g] - b4[:, :g]).norm() / b4[:, :g].norm()).item()
    print("first-step gradient blob, atomic vs ordered accumulation: rel-L2", rel, "losses", l3, l4)
    assert rel <= 1e-5 and abs(l3[0] - l4[0]) <= 1e-6 * max(1., abs(l4[0]))
