"""Edge cases of the C-ABI entry points on the device: empty inputs, ragged row counts that do not
fill a 128-row tile or a tile pair, sample counts that are not multiples of 32, maximum merge size."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.helpers import load_model_params, record

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _K():
    import cv_nerf_b200
    return cv_nerf_b200.kernels


def _net(seed=0):
    from cv_nerf_b200.model import Model
    p, _ = O.init_field_params(seed, 1.0, 5.0)
    return p, load_model_params(Model(), p).to(DEV)


def test_empty_inputs_are_no_ops():
    K = _K()
    p, net = _net()
    e = lambda *s: torch.empty(s, device=DEV)
    assert K.pack_rays(8, 8, 10., rays_o=e(0, 3), rays_d=e(0, 3), ndc=False).shape == (0, 11)
    rays = e(0, 11)
    assert K.sample_coarse(rays, 64).shape == (0, 64)
    rgb, w = K.composite_fwd(e(0, 64, 4), e(0, 64), e(0, 3))
    assert rgb.shape == (0, 3) and w.shape == (0, 64)
    assert K.resample_merge(e(0, 64), e(0, 64), e(0, 128)).shape == (0, 192)
    assert K.sample_pdf(e(0, 63), e(0, 62), e(0, 128)).shape == (0, 128)
    packed = net.packed()
    vt = K.viewdir_term(packed, e(0, 3))
    assert K.mlp_fwd(packed, K.IN_POINTS, e(0, 3), None, 0, 1, e(1, 128), 1).shape == (0, 4)
    assert K.mlp_fwd(packed, K.IN_POINTS, e(0, 3), None, 0, 1, e(1, 128), 1, host_tail=net.host_tail()).shape == (0, 4)
    assert K.freq_encode(e(0, 3), 10).shape == (0, 63)
    assert K.to_byte(e(0, 3)).shape == (0, 3)
    assert K.composite_maps(e(0, 8), e(0, 8)).shape == (0, 3)
    torch.cuda.synchronize()


@pytest.mark.parametrize("n,S_c,n_fine", [(1, 64, 128), (37, 64, 128), (5, 33, 7), (3, 3, 1), (2, 64, 192)])
def test_ragged_render_matches_oracle(n, S_c, n_fine):
    """Ray counts / sample counts that leave tiles, warps and tile pairs partly empty."""
    from cv_nerf_b200 import main as M
    cp, coarse = _net(1)
    fp, fine = _net(2)
    g = torch.Generator().manual_seed(n * 1000 + S_c)
    o = torch.randn(n, 3, generator=g) * .2 + torch.tensor([0., 0., 4.])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * torch.tensor([1., 1., -1.])
    u = torch.rand(n, n_fine, generator=g)
    kw = dict(white_bkg=True, ndc=False, near=2., far=6.)
    with torch.no_grad():
        rgb, extras = M.render(16, 16, 20., rays=torch.stack([o, d]).to(DEV), draws=M.RenderDraws(u=u), coarse_model=coarse,
                               fine_model=fine, n_coarse_samples=S_c, n_fine_samples=n_fine, **kw)
    ref = O.render_image(16, 16, 20., cp, fp, rays=(o, d), n_coarse=S_c, n_fine=n_fine, draws=O.RenderDraws(u=u), **kw)
    assert rgb.shape == (n, 3)
    assert (rgb.cpu() - ref["rgb_map"]).abs().max().item() <= 1e-2
    assert (extras["rgb_c"].cpu() - ref["rgb_c"]).abs().max().item() <= 1e-2


def test_backward_with_partial_tiles_and_single_row():
    """Gradients for row counts of 1, 127 and 129 (one row, just under / over a tile)."""
    K = _K()
    p, net = _net(3)
    for rows in (1, 127, 129):
        torch.manual_seed(rows)
        x = torch.cat([O.freq_encode(torch.rand(rows, 3) * 2 - 1, 10),
                       O.freq_encode(torch.nn.functional.normalize(torch.randn(rows, 3), dim=-1), 4)], -1)
        gr = torch.randn(rows, 4)
        q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        O.field_mlp(q, x).backward(gr)
        net.zero_grad()
        net(x.to(DEV)).backward(gr.to(DEV))
        for name, prm in net.named_parameters():
            ref = q[name].grad
            err = (prm.grad.cpu() - ref).norm().item() / max(ref.norm().item(), 1e-12)
            record("grad_partial_tiles", dict(rows=rows, tensor=name, rel_l2=err))
            assert err <= 0.2, (rows, name, err)        # few rows: single ReLU-mask flips weigh more


def test_argument_errors_are_reported():
    import cv_nerf_b200
    from cv_nerf_b200 import _lib
    lib = _lib.load()
    rc = lib.nerf_sample_coarse(None, 5, 64, None, None, None)
    assert rc < 0 and b"nerf_sample_coarse" in lib.nerf_b200_last_error()
    rc = lib.nerf_resample_merge(None, None, None, 1, 200, 100, None, None)    # S + m > 256
    assert rc != 0
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        _K().composite_fwd(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))   # CPU tensors
