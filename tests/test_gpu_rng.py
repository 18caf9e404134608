"""In-kernel random draws (csrc/rng.cuh) on the device: the CUDA Philox equals the numpy restatement
(which is pinned to the Random123 known answers), the consuming kernels draw exactly the numbers
nerf_rng_fill writes out, and a render keyed by an Rng equals the render with those numbers injected
-- which is the path the parity tests check against the oracle and the reference fixtures."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import philox_ref as P
from tests.helpers import load_model_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _K():
    import cv_nerf_b200
    return cv_nerf_b200.kernels


@pytest.mark.parametrize("seed,stream,ray0,n,cols", [(0, 0, 0, 33, 64), (0x1234567890ABCDEF, 1, 5_000_000_000, 17, 128),
                                                     (42, 3, 7, 9, 191), (42, 2, 0, 5, 3)])
def test_rng_fill_equals_numpy_philox(seed, stream, ray0, n, cols):
    K = _K()
    rng = K.Rng(seed, ray0)
    u = K.rng_fill("uniform", rng, stream, n, cols, DEV).cpu().numpy()
    assert np.array_equal(u, P.uniforms(seed, stream, ray0, n, cols)), "uniform draws differ from Philox4x32-10"
    z = K.rng_fill("normal", rng, stream, n, cols, DEV).cpu().numpy()
    np.testing.assert_allclose(z, P.normals(seed, stream, ray0, n, cols), rtol=0, atol=5e-6)


def _scene(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g) * .2 + torch.tensor([0., 0., 4.])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * torch.tensor([1., 1., -1.])
    return torch.stack([o, d]).to(DEV)


def _nets():
    from cv_nerf_b200.model import Model
    cp, fp = O.init_field_params(1, 1.0, 5.0)
    return load_model_params(Model(), cp).to(DEV), load_model_params(Model(), fp).to(DEV)


@pytest.mark.parametrize("perturb,noise,S_c,n_fine", [(0., 0., 64, 128), (1., 1., 64, 128), (1., 0.5, 33, 7), (1., 1., 130, 100)])
def test_render_with_rng_equals_render_with_the_same_draws_injected(perturb, noise, S_c, n_fine):
    from cv_nerf_b200 import main as M
    K = _K()
    coarse, fine = _nets()
    n = 203
    rays = _scene(n)
    rng = K.Rng(99, 1000)
    kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=S_c, n_fine_samples=n_fine, white_bkg=False,
              ndc=False, near=2., far=6., perturb=perturb, noise=noise)
    with torch.no_grad():
        a, ea = M.render(16, 16, 20., rays=rays, rng=rng, extras=True, **kw)
        draws = M.RenderDraws(u=K.rng_fill("uniform", rng, K.RNG_U, n, n_fine, DEV))
        if perturb > 0:
            draws.t_rand = K.rng_fill("uniform", rng, K.RNG_T_RAND, n, S_c, DEV)
        if noise > 0:
            draws.noise_c = K.rng_fill("normal", rng, K.RNG_NOISE_C, n, S_c, DEV)
            draws.noise_f = K.rng_fill("normal", rng, K.RNG_NOISE_F, n, S_c + n_fine, DEV)
        b, eb = M.render(16, 16, 20., rays=rays, draws=draws, extras=True, **kw)
    for key in ("z_c", "z_f"):
        assert torch.equal(ea[key], eb[key]), key
    assert torch.equal(a, b) and torch.equal(ea["rgb_c"], eb["rgb_c"])
    # chunked calls draw what the whole call draws (keys carry the global ray index).  The pixels are bit-identical
    # when the split falls on a 128-row tile boundary of both passes (the field kernel sums a row's K chunks in
    # the order given by the parity of its global tile, include/nerf_b200.h); otherwise equal to fp32 summation order
    with torch.no_grad():
        parts = [M.render(16, 16, 20., rays=rays[:, i:j], rng=rng.shifted(i), **kw)[0] for i, j in ((0, 50), (50, 203))]
    got = torch.cat(parts, 0)
    if (50 * S_c) % 128 == 0 and (50 * (S_c + n_fine)) % 128 == 0:
        assert torch.equal(got, a)
    else:
        assert (got - a).abs().max().item() <= 2e-6


def test_row_sharded_frame_with_rng_equals_whole_frame_and_seed_follows_torch():
    from cv_nerf_b200 import main as M
    coarse, fine = _nets()
    pose = O.lego_pose(20., -30., 4.)[:3, :4].to(DEV)
    kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=True, ndc=False,
              near=2., far=6., perturb=1., noise=0.3)
    rng = _K().Rng(5)
    with torch.no_grad():
        whole, _ = M.render(24, 32, 40., c2w=pose, rng=rng, **kw)
        parts = [M.render(24, 32, 40., c2w=pose, rows=r, rng=rng, **kw)[0] for r in ((0, 7), (7, 16), (16, 24))]
        assert torch.equal(torch.cat(parts, 0), whole)
        torch.manual_seed(3); x = M.render(24, 32, 40., c2w=pose, **kw)[0]
        torch.manual_seed(3); y = M.render(24, 32, 40., c2w=pose, **kw)[0]
        z = M.render(24, 32, 40., c2w=pose, **kw)[0]
    assert torch.equal(x, y) and not torch.equal(x, z)


def test_composite_backward_regenerates_the_forward_noise():
    K = _K()
    torch.manual_seed(0)
    n, s = 77, 192
    raw = torch.randn(n, s, 4, device=DEV)
    z = torch.sort(torch.rand(n, s, device=DEV) * 4 + 2, -1).values
    d = torch.randn(n, 3, device=DEV)
    g = torch.randn(n, 3, device=DEV)
    key = K.RngNoise(0.7, K.Rng(11, 300), K.RNG_NOISE_F)
    noise = K.rng_fill("normal", key.rng, key.stream, n, s, DEV) * 0.7
    rgb_a, w_a = K.composite_fwd(raw, z, d, key, False)
    rgb_b, w_b = K.composite_fwd(raw, z, d, noise, False)
    assert torch.equal(rgb_a, rgb_b) and torch.equal(w_a, w_b)
    assert torch.equal(K.composite_bwd(raw, z, d, key, False, g), K.composite_bwd(raw, z, d, noise, False, g))


def test_resample_merge_sorts_and_handles_unsorted_coarse_depths():
    """z_f == sort(cat(z_c, samples)) bit for bit, for ascending, descending and shuffled coarse rows."""
    K = _K()
    g = torch.Generator().manual_seed(4)
    for S, m in ((64, 128), (128, 128), (33, 7), (3, 1), (100, 28), (130, 100)):
        n = 301
        z = torch.sort(torch.rand(n, S, generator=g) * 4 + 2, -1).values
        z[1] = z[1].flip(0)
        z[2] = z[2][torch.randperm(S, generator=g)]
        z[3, : S // 2] = z[3, S // 2]                     # ties
        w = torch.rand(n, S, generator=g) ** 3
        u = torch.rand(n, m, generator=g)
        got = K.resample_merge(z.to(DEV), w.to(DEV), u.to(DEV)).cpu()
        mids = .5 * (z[:, 1:] + z[:, :-1])
        smp = K.sample_pdf(mids.to(DEV), w[:, 1:-1].contiguous().to(DEV), u.to(DEV)).cpu()
        want = torch.sort(torch.cat([z, smp], -1), -1).values
        assert torch.equal(got, want), (S, m, (got - want).abs().max())
        # the samples themselves against the oracle on the well-formed rows (rows 1..3 have unsorted or
        # tied bins; only their merge is checked).  The reference's rule "cdf span < 1e-5 -> divide by 1"
        # (utils.py:45-46) is a step: a bin whose span sits at the threshold can take the other branch when
        # two cumsum orders differ in the last bit, which moves that one sample by up to the bin width --
        # so the bound is on all but a 1e-3 fraction of the samples, and a loose one on the rest.
        ref = O.inverse_cdf_sample(mids, w[:, 1:-1], u)
        diff = (smp[4:] - ref[4:]).abs()
        assert (diff > 5e-5).float().mean().item() <= 1e-3 and diff.max().item() <= 0.5, (S, m, diff.max().item())


@pytest.mark.parametrize("ndc,perturb,noise", [(False, 0., 0.), (True, 1., 1.)])
def test_fused_whole_chain_entry_equals_the_launch_by_launch_path(ndc, perturb, noise, monkeypatch):
    """nerf_render_fused (one C call sequencing the chain) returns the pixels of the Python-sequenced
    path bit for bit, for a pose-driven frame, a row shard of it and a caller-provided ray batch."""
    from cv_nerf_b200 import main as M
    coarse, fine = _nets()
    pose = O.lego_pose(20., -30., 4.)[:3, :4].to(DEV)
    near, far = (0., 1.) if ndc else (2., 6.)
    kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=not ndc, ndc=ndc,
              near=near, far=far, perturb=perturb, noise=noise)
    rng = _K().Rng(77)
    rays = _scene(301)
    with torch.no_grad():
        assert M._fused_inference(kw, None)
        a, ea = M.render(24, 32, 40., c2w=pose, rng=rng, **kw)
        s, _ = M.render(24, 32, 40., c2w=pose, rows=(5, 17), rng=rng, **kw)
        r, er = M.render(24, 32, 40., rays=rays, rng=rng, **kw)
        monkeypatch.setenv("NERF_B200_FUSED_RENDER", "0")
        assert not M._fused_inference(kw, None)
        b, eb = M.render(24, 32, 40., c2w=pose, rng=rng, **kw)
        r2, er2 = M.render(24, 32, 40., rays=rays, rng=rng, **kw)
    assert a.shape == (24, 32, 3) and torch.equal(a, b) and torch.equal(ea["rgb_c"], eb["rgb_c"])
    assert torch.equal(s, b[5:17])
    assert r.shape == (301, 3) and torch.equal(r, r2) and torch.equal(er["rgb_c"], er2["rgb_c"])
    # with autograd recording the launch-by-launch path is taken (activation records are needed)
    monkeypatch.delenv("NERF_B200_FUSED_RENDER")
    assert not M._fused_inference(kw, None)
