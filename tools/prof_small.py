"""One 800x800 frame (BASELINE.json configs[1]) and one 4096-ray train step (configs[3]) for an ncu
capture of the memory-bound kernels around the field network:

    ncu --set full --clock-control none -k regex:'composite|resample|pack_rays|sample_coarse|viewdir|mse_loss|train_rays' \
        -c 40 -o gpurun_out/small python tools/prof_small.py

Summaries are kept under profiles/ (r02_small_kernels_ncu.txt)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200 import main as M  # noqa: E402
from cv_nerf_b200.data_helpers import pose_spherical  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402
from cv_nerf_b200.train import TrainStep  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 800
dev = torch.device("cuda:0")
torch.manual_seed(0)
coarse, fine = Model().to(dev), Model().to(dev)
pose = pose_spherical(-180., -30., 4.)[:3, :4].to(dev)
kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=True, ndc=False,
          near=2., far=6.)
with torch.no_grad():
    rgb, _ = M.render(size, size, 1111.111 * size / 800., c2w=pose, **kw)
torch.cuda.synchronize()
ts = TrainStep(coarse, fine, height=400, width=400, focal=555.5555, n_rays=4096, perturb=1., noise=0.,
               white_bkg=True, ndc=False, near=2., far=6., lr=5e-4, lr_decay=500, seed=1)
image = torch.rand(400, 400, 3, device=dev)
loss = ts.step(image, pose)
torch.cuda.synchronize()
print("ok", float(rgb.mean()), float(loss))
