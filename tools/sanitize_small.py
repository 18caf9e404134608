"""Small end-to-end workload (one 16x16 render, one 96-ray train step, odd sizes) for
`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py` (one tool per call)."""
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

dev = torch.device("cuda:0")
torch.manual_seed(0)
coarse, fine = Model().to(dev), Model().to(dev)
pose = pose_spherical(-180., -30., 4.)[:3, :4].to(dev)
kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=True, ndc=False,
          near=2., far=6.)
with torch.no_grad():
    rgb, _ = M.render(13, 17, 20., c2w=pose, **kw)          # 221 rays: partial tiles and tile pairs
ts = TrainStep(coarse, fine, height=32, width=32, focal=40., n_rays=96, perturb=1., noise=1., white_bkg=False,
               ndc=True, near=0., far=1.)
image = torch.rand(32, 32, 3, device=dev)
loss = ts.step(image, pose)
torch.cuda.synchronize()
print("ok", float(rgb.mean()), float(loss))
