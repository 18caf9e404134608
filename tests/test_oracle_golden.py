"""The oracle (oracle/nerf_oracle.py) replayed against fixtures recorded from the real
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O


def _sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _focal(g, prefix=""):
    f = float(g[prefix + "hwf"][2])
    return np.float32(f) if bool(g[prefix + "f_is_f32"]) else f


@pytest.mark.parametrize("case", ["lego400", "lego800", "fern"])
def test_ray_grid_bit_exact(golden_dir, case):
    g = _load(golden_dir, "rays.npz")
    h, w = int(g[f"{case}_hwf"][0]), int(g[f"{case}_hwf"][1])
    f = _focal(g, case + "_")
    o, d = O.ray_grid(h, w, f, torch.from_numpy(g[f"{case}_pose"]))
    assert _sha(d) == str(g[f"{case}_d_sha"])
    assert _sha(o.contiguous()) == str(g[f"{case}_o_sha"])
    assert np.array_equal(d[::37, ::41].numpy(), g[f"{case}_d_sub"])


def test_ndc_bit_exact(golden_dir):
    g = _load(golden_dir, "rays.npz")
    h, w = int(g["fern_hwf"][0]), int(g["fern_hwf"][1])
    f = _focal(g, "fern_")
    o, d = O.ray_grid(h, w, f, torch.from_numpy(g["fern_pose"]))
    on, dn = O.ndc_warp(h, w, f, 1., o, d)
    assert _sha(on) == str(g["fern_ndc_o_sha"])
    assert _sha(dn) == str(g["fern_ndc_d_sha"])


@pytest.mark.parametrize("tag", ["c", "f"])
def test_composite_matches(golden_dir, tag):
    g = _load(golden_dir, "units.npz")
    raw, z, d = (torch.from_numpy(g[f"comp_{tag}_{k}"]) for k in ("raw", "z", "d"))
    rgb, w = O.composite(raw, z, d, None, True)
    assert np.array_equal(rgb.numpy(), g[f"comp_{tag}_rgb_white"])
    assert np.array_equal(w.numpy(), g[f"comp_{tag}_w_white"])
    nd = torch.from_numpy(g[f"comp_{tag}_noise_draw"]) * 0.7
    rgb, w = O.composite(raw, z, d, nd, False)
    assert np.array_equal(rgb.numpy(), g[f"comp_{tag}_rgb_noise"])
    assert np.array_equal(w.numpy(), g[f"comp_{tag}_w_noise"])


def test_inverse_cdf_matches(golden_dir):
    g = _load(golden_dir, "units.npz")
    s = O.inverse_cdf_sample(*(torch.from_numpy(g[k]) for k in ("pdf_bins", "pdf_w", "pdf_u")))
    assert np.array_equal(s.numpy(), g["pdf_samples"])


def test_freq_encode_matches(golden_dir):
    g = _load(golden_dir, "units.npz")
    x = torch.from_numpy(g["enc_x"])
    assert np.array_equal(O.freq_encode(x, 10).numpy(), g["enc10"])
    assert np.array_equal(O.freq_encode(x, 4).numpy(), g["enc4"])


def _replay(g, extras=False):
    h, w = int(g["hwf"][0]), int(g["hwf"][1])
    f = _focal(g)
    sb = None if np.isnan(g["sigma_bias"]) else float(g["sigma_bias"])
    coarse, fine = O.init_field_params(int(g["seed"]), sb, float(g["sigma_gain"]))
    draws = O.RenderDraws(u=torch.from_numpy(g["u"]))
    noise = float(g["noise"])
    if bool(g["train"]):
        draws.t_rand = torch.from_numpy(g["t_rand"])
        if noise > 0:
            draws.noise_c = torch.from_numpy(g["noise_c"]) * noise
            draws.noise_f = torch.from_numpy(g["noise_f"]) * noise
    rays = (torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]))
    return coarse, fine, O.render_image(h, w, f, coarse, fine, rays=rays, ndc=bool(g["ndc"]),
                                        near=float(g["near"]), far=float(g["far"]), draws=draws,
                                        white_bkg=bool(g["white_bkg"]), extras=extras)


RENDER_FIXTURES = ["lego_test", "lego_test_stock", "fern_test", "lego_train", "fern_train",
                   # round 2 (make_golden_r02.py): headline 800x800 config, the reference's skull spiral poses
                   "lego800_test", "skull_f000_test", "skull_f030_test", "skull_f060_test", "skull_f090_test"]


@pytest.mark.parametrize("name", RENDER_FIXTURES)
def test_render_matches_reference(golden_dir, name):
    g = _load(golden_dir, f"render_{name}.npz")
    sb = g["sigma_bias"]
    coarse, fine, out = _replay(g)
    h = hashlib.sha256()
    for p in (coarse, fine):
        for k, v in p.items():
            h.update(k.encode()); h.update(v.numpy().tobytes())
    assert h.hexdigest() == str(g["weights_sha"]), "init order / RNG differs from the reference"
    # same torch build => bit-exact; allow MKL code-path differences across hosts
    np.testing.assert_allclose(out["rgb_map"].numpy(), g["rgb_map"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["rgb_c"].numpy(), g["rgb_c"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["lego_train", "fern_train"])
def test_gradients_match_reference(golden_dir, name):
    g = _load(golden_dir, f"render_{name}.npz")
    h, w = int(g["hwf"][0]), int(g["hwf"][1])
    sb = float(g["sigma_bias"])
    coarse, fine = O.init_field_params(int(g["seed"]), sb, float(g["sigma_gain"]))
    for p in (coarse, fine):
        for v in p.values():
            v.requires_grad_(True)
    noise = float(g["noise"])
    draws = O.RenderDraws(u=torch.from_numpy(g["u"]), t_rand=torch.from_numpy(g["t_rand"]))
    if noise > 0:
        draws.noise_c = torch.from_numpy(g["noise_c"]) * noise
        draws.noise_f = torch.from_numpy(g["noise_f"]) * noise
    rays = (torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]))
    out = O.render_image(h, w, _focal(g), coarse, fine, rays=rays, ndc=bool(g["ndc"]),
                         near=float(g["near"]), far=float(g["far"]), draws=draws,
                         white_bkg=bool(g["white_bkg"]))
    loss = O.train_loss(out["rgb_map"], out["rgb_c"], torch.from_numpy(g["target"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    for tag, p in (("coarse", coarse), ("fine", fine)):
        for k, v in p.items():
            gn = float(g[f"gnorm/{tag}.{k}"])
            assert abs(v.grad.norm().item() - gn) <= 1e-5 * max(gn, 1e-3), (tag, k)
            np.testing.assert_allclose(v.grad.reshape(-1)[:96].numpy(), g[f"ghead/{tag}.{k}"],
                                       rtol=1e-4, atol=1e-7)
