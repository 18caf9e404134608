"""A few train iterations (BASELINE.json configs[3] shape: 4096 rays, 64+128 samples) for ncu captures
of the training kernels.  Usage: python tools/prof_train.py [steps] [n_rays]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200.data_helpers import pose_spherical  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402
from cv_nerf_b200.train import TrainStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
torch.manual_seed(0)
coarse, fine = Model().to(dev), Model().to(dev)
ts = TrainStep(coarse, fine, height=400, width=400, focal=555.5555, n_rays=n_rays, perturb=1., noise=0.,
               white_bkg=True, ndc=False, near=2., far=6., lr=5e-4, lr_decay=500, seed=1)
image = torch.rand(400, 400, 3, device=dev)
pose = pose_spherical(-180., -30., 4.)[:3, :4].to(dev)
for i in range(steps):
    loss = ts.step(image, pose)
torch.cuda.synchronize()
print("loss", loss.item())
