"""CPU-only checks: the C-ABI library loads and exports every symbol include/nerf_b200.h declares,
host-side logic (config reader, Model surface, constant folding) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import cv_nerf_b200
    from cv_nerf_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nerf_b200.h but not exported"
    assert sorted(_lib.public_symbols()) == declared, "ctypes table and header disagree"
    assert lib.nerf_b200_abi_version() == 1
    assert lib.nerf_packed_model_bytes() == 64 * 16384 + 4 * (9 * 256 + 256 + 4 + 384 + 4 + 128 * 28 + 128)


def test_no_cpu_fallback():
    import cv_nerf_b200
    from cv_nerf_b200 import main as M, utils as U, model as MD
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        M.compute_rays(8, 8, 10., torch.eye(4)[:3])
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        U.inv_transform_sampling(torch.rand(4, 63), torch.rand(4, 62), 16)
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        MD.Model()(torch.rand(3, 90))
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        MD.FreqEmbedding(10).embed(torch.rand(3, 3))


def test_gradient_wrappers_refuse_host_buffers():
    """the split gradient launches (heads / dW / view columns) check every buffer they hand to the C ABI"""
    import cv_nerf_b200
    from cv_nerf_b200 import kernels as K
    act = torch.zeros(64, dtype=torch.uint8)
    blob = torch.zeros(16)
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        K.mlp_bwd_heads(act, torch.zeros(4, 4), 4, blob)
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        K.mlp_bwd_dw(act, act, 4, blob)
    with pytest.raises(cv_nerf_b200.NerfB200Error):
        K.viewdir_term_bwd(act, 4, torch.zeros(1, 11), 4, False, blob)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cv-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|nerf_oracle|oracle/", src, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_model_surface_matches_reference_layout():
    from cv_nerf_b200.model import Model, FreqEmbedding
    from oracle import nerf_oracle as O
    torch.manual_seed(0)
    a, b = Model(), Model()
    coarse, fine = O.init_field_params(0)
    sd = a.state_dict()
    assert list(sd.keys()) == [f"{n}.{k}" for n in ("l1", "l2", "l3", "l4", "l5", "l6", "l7", "l8", "l9",
                                                     "l_alpha", "l10", "l11") for k in ("weight", "bias")]
    for k, v in sd.items():
        assert tuple(v.shape) == (O.LAYER_SHAPES[k.split(".")[0]] if k.endswith("weight")
                                  else (O.LAYER_SHAPES[k.split(".")[0]][0],))
        assert torch.equal(v, coarse[k]), "default init must match the reference's RNG order"
    for k, v in b.state_dict().items():
        assert torch.equal(v, fine[k])
    assert sum(p.numel() for p in a.parameters()) == 595844
    assert FreqEmbedding(10).out_dim == 63 and FreqEmbedding(4).out_dim == 27


def test_load_config_reads_reference_flag_files(tmp_path):
    from cv_nerf_b200.main import load_config
    p = tmp_path / "lego.txt"
    p.write_text("name = blender_paper_lego\ndtype = blender\nwhite_bkg = True\nlr_decay = 500\n"
                 "n_coarse_samples = 64\nn_fine_samples = 128\nn_rays = 1024\nprecrop_iters = 500\n"
                 "precrop_frac = 0.5\nhalf_res = True\nexpname = leftover\n")
    a = load_config(str(p))
    assert a.dtype == "blender" and a.white_bkg is True and a.n_fine_samples == 128 and a.n_rays == 1024
    assert a.precrop_frac == 0.5 and a.half_res is True and a.chunk == 32768 and a.netchunk == 65536
    assert a.perturb == 1. and a.noise == 0. and a.lr == 5e-4 and not hasattr(a, "expname")
    d = load_config()
    assert d.n_fine_samples == 0 and d.dtype == "llff" and d.n_rays == 4096


def test_unit_linspace_emulation_matches_torch():
    """csrc/rays.cu reproduces torch.linspace(0,1,S) as step*i below the midpoint and
    fma(-step, S-1-i, 1) above it; keep that in sync with the installed torch."""
    for S in (64, 33, 128, 7, 2, 192):
        step = np.float32(1.0) / np.float32(S - 1)
        i = np.arange(S)
        lo = (np.float64(step) * i).astype(np.float32)
        hi = (1.0 - np.float64(step) * (S - 1 - i)).astype(np.float32)
        emul = np.where(i < S // 2, lo, hi)
        assert np.array_equal(emul, torch.linspace(0., 1., S).numpy()), S


def test_ndc_constant_folding():
    from cv_nerf_b200.kernels import _ndc_consts
    for focal in (np.float32(407.5657), 555.5555155968841, np.float64(404.72647)):
        cw, ch = _ndc_consts(378, 504, focal)
        want_w = torch.tensor(-1. / (504 / (2. * focal))).float().item()
        want_h = torch.tensor(-1. / (378 / (2. * focal))).float().item()
        assert cw == want_w and ch == want_h


def test_checkpoint_round_trip_keeps_reference_keys(tmp_path):
    """state_dict keys/shapes are the reference's (model.py:57-71), so its checkpoints load; the
    save/load helpers round-trip both networks and the step counter (host-only, no kernels)."""
    import cv_nerf_b200
    from cv_nerf_b200.model import Model
    from cv_nerf_b200.train import load_checkpoint, save_checkpoint
    from oracle import nerf_oracle as O
    torch.manual_seed(0)
    a, b = Model(), Model()
    sd = a.state_dict()
    assert set(sd) == {f"{n}.{k}" for n in O.LAYER_NAMES for k in ("weight", "bias")}
    for n in O.LAYER_NAMES:
        assert tuple(sd[n + ".weight"].shape) == O.LAYER_SHAPES[n]
    path = tmp_path / "000123.pt"
    save_checkpoint(str(path), 123, a, b)
    c, d = Model(), Model()
    assert load_checkpoint(str(path), c, d) == 123
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), c.state_dict().values()))
    assert all(torch.equal(x, y) for x, y in zip(b.state_dict().values(), d.state_dict().values()))


def test_write_png_round_trips(tmp_path):
    """render_full writes '{:03d}.png' like the reference (main.py:117-120); decode the file by hand."""
    import struct
    import zlib
    import numpy as np
    from cv_nerf_b200.main import write_png
    img = (np.arange(7 * 5 * 3, dtype=np.uint32) * 37 % 256).astype(np.uint8).reshape(7, 5, 3)
    path = tmp_path / "000.png"
    write_png(str(path), img)
    data = path.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(data):
        (n,), tag = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xffffffff
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        if tag == b"IDAT":
            idat += body
        pos += 12 + n
    assert hdr == (5, 7, 8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(7, 1 + 5 * 3)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(7, 5, 3), img)
