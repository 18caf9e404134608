"""A/B timing of the gradient-kernel schedules of TrainStep.forward_backward (NERF_B200_BWD_SCHED):
0 serial on one stream, 1 small kernels beside dW, 2 heads under the dZ chains, view columns beside dW.
Usage: python tools/time_bwd_schedules.py [n_rays] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200.data_helpers import pose_spherical  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402
from cv_nerf_b200.train import TrainStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
torch.manual_seed(0)
coarse, fine = Model().to(dev), Model().to(dev)
ts = TrainStep(coarse, fine, height=400, width=400, focal=555.5555, n_rays=n, perturb=1., noise=0., white_bkg=True,
               ndc=False, near=2., far=6., seed=1)
image = torch.rand(400, 400, 3, device=dev)
pose = pose_spherical(-180., -30., 4.)[:3, :4].to(dev)
configs = [(0, 0), (1, 0), (2, 0)]
for rnd in range(2):
    for sched, split in configs:
        os.environ["NERF_B200_BWD_SCHED"] = str(sched)
        os.environ["NERF_B200_BWD_SPLIT"] = str(split)
        for _ in range(3):
            ts.step(image, pose)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ts.step(image, pose)
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} schedule {sched} split {split}: {e0.elapsed_time(e1) / reps:.3f} ms per train step ({n} rays)")
