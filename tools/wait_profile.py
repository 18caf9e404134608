"""Sampled wait profile of the production field kernel's MMA thread (debug variant 18): cycles spent
waiting for weight slots per (layer, chunk) and for the A operand per (layer, sub-tile), averaged per
sampled tile pair over all CTAs.  Usage: python tools/wait_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200 import _lib, kernels as K  # noqa: E402
from tools.gpu_diag import packed_model, DEV  # noqa: E402

lib = _lib.load()
if not _lib.has_experiments():
    sys.exit("needs the experiments build: make -C cv-nerf_b200/csrc experiments && "
             "NERF_B200_LIB=$PWD/cv-nerf_b200/libnerf_b200_exp.so python " + sys.argv[0])
VARIANT = int(sys.argv[1]) if len(sys.argv) > 1 else 18     # 18: mlp_fwd_kernel, 31: the CTA-pair kernel
p, packed = packed_model()
n_rays, S = 160000, 192
rays = torch.zeros(n_rays, 11, device=DEV)
rays[:, 0:3] = torch.randn(n_rays, 3, device=DEV) * .3
rays[:, 3:6] = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV), dim=-1)
rays[:, 6], rays[:, 7] = 2., 6.
rays[:, 8:11] = rays[:, 3:6]
z = K.sample_coarse(rays, S)
vt = K.viewdir_term(packed, rays)
raw = torch.empty(n_rays * S, 4, device=DEV)
buf = torch.zeros(148 * 8 + 4 * 1024, dtype=torch.int64, device=DEV)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    buf.zero_()
    rc = lib.nerf_mlp_fwd_stats(packed.data_ptr(), rays.data_ptr(), z.data_ptr(), n_rays * S, S, vt.data_ptr(),
                                raw.data_ptr(), 1000 + VARIANT, buf.data_ptr(), st)
    torch.cuda.synchronize()
    assert rc == 0, lib.nerf_b200_last_error()
tr = buf[148 * 8:].cpu().tolist()
pairs = max(tr[100], 1)
c5 = buf[:148 * 8].view(148, 8)[:, 5].double()
cyc = c5[c5 > 0].mean().item()
print(f"sampled tile pairs: {pairs}; kernel cycles per CTA {cyc:.3e}")
print("weight-slot wait, cycles per tile pair (both sub-tiles), by layer and chunk:")
tot_w = 0
for l in range(9):
    row = [tr[l * 5 + j] / pairs for j in range(5)]
    tot_w += sum(row)
    print(f"  layer {l}: " + "  ".join(f"{v:7.0f}" for v in row))
print("A-operand wait, cycles per tile pair, by layer (X, Y):")
tot_a = 0
for l in range(9):
    a = [tr[64 + l * 2 + g] / pairs for g in range(2)]
    tot_a += sum(a)
    print(f"  layer {l}: {a[0]:7.0f} {a[1]:7.0f}")
print(f"per tile pair: weight-slot waits {tot_w:.0f}, A-operand waits {tot_a:.0f} cycles")
print("epilogue of the first warp of each group, accumulator complete -> A operand handed over, cycles (X, Y):")
for l in range(8):
    print(f"  layer {l}: {tr[104 + l * 2] / pairs:7.0f} {tr[105 + l * 2] / pairs:7.0f}")
print(f"  rgb epilogue + next tile's PE: {tr[101] / pairs:7.0f} {tr[102] / pairs:7.0f}")
if VARIANT == 31:
    n = 16 * pairs
    print(f"hidden epilogue of the first warp (mean over 8 layers x 2 groups): body {tr[130] / n:.0f}, proxy + tcgen05 fences "
          f"{tr[131] / n:.0f}, warp sync + remote arrive {tr[132] / n:.0f} cycles")
