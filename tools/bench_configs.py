"""Measurements for the BASELINE.json configs that bench.py does not headline:
  configs[2]  fern 378x504 NDC: one-frame render (rays/s) and train step (ms, 4096 rays)
  configs[4]  skull 504x378 NDC: 120-frame spiral video (frames/s, rays/s), frame-parallel over ranks
Prints one JSON line per measurement (rank 0).  Poses: tests/golden/skull_spiral.npz (derived with the
reference's pose math); fern poses are not in the reference tree, so the fern shape is rendered from
the recentred skull training poses.  Synthetic weights (torch seed 0).

    python tools/bench_configs.py                # 1 GPU
    torchrun --nproc-per-node N tools/bench_configs.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200 import main as M  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402
from cv_nerf_b200.train import TrainStep  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))

g = np.load(os.path.join(ROOT, "tests", "golden", "skull_spiral.npz"))
torch.manual_seed(0)
coarse, fine = Model().to(dev), Model().to(dev)
kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=False,
          ndc=True, near=0., far=1., perturb=False, noise=0.)


def emit(**d):
    if rank == 0:
        print(json.dumps(d), flush=True)


def timed(fn, reps):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / reps


# ---- configs[2]: fern shape, NDC, single frame + train step -----------------------------------
H, W, F = 378, 504, np.float32(407.5657)
fern_pose = torch.from_numpy(g["train_poses"][3]).to(dev)
from cv_nerf_b200 import parallel as P  # noqa: E402
bounds = P.row_bounds(H, world)


def fern_frame():
    with torch.no_grad():
        rgb, _ = M.render(H, W, F, c2w=fern_pose, rows=(bounds[rank], bounds[rank + 1]), **kw)
        P.all_gather_rows(rgb, H)


sec = timed(fern_frame, 5)
emit(config="fern 378x504 NDC render, 64+128 samples, rows sharded over the GPUs", metric="rendered rays/sec",
     value=H * W / sec, unit="rays/s", ms_per_frame=sec * 1e3, n_gpus=world)
ts = TrainStep(coarse, fine, height=H, width=W, focal=F, n_rays=4096, perturb=1., noise=1., white_bkg=False, ndc=True,
               near=0., far=1., lr=5e-4, lr_decay=250, seed=rank)
image = torch.rand(H, W, 3, device=dev)
sec = timed(lambda: ts.step(image, fern_pose), 20)
emit(config="fern train step (configs/fern.txt: perturb 1, noise 1), 4096 rays per GPU", metric="train step ms",
     value=sec * 1e3, unit="ms", n_gpus=world, rays_per_s_all_gpus=4096 * world / sec)

# ---- configs[4]: skull 120-frame spiral video ------------------------------------------------------
H, W, F = int(g["hwf"][0]), int(g["hwf"][1]), np.float32(g["hwf"][2])
poses = [torch.from_numpy(p).to(dev) for p in g["render_poses"]]
torch.manual_seed(0)
coarse2, fine2 = Model().to(dev), Model().to(dev)
kw.update(coarse_model=coarse2, fine_model=fine2)
# sweep over the reference's `chunk` flag (configs/skull.txt: 32768; the inference path merges chunks up to
# 2^20 rays per launch, so the flag only matters through NERF_B200_FUSED_RENDER=0 + autograd) and the frame format
for as_bytes, chunk in ((False, 32768), (True, 32768), (True, 1 << 20)):
    sec = timed(lambda: M.render_full(poses, [H, W, F], chunk, kw, as_bytes=as_bytes, verbose=False), 1)
    emit(config=f"skull 504x378 NDC 120-frame spiral video ({'uint8' if as_bytes else 'float32'} frames to host)",
         metric="video render", value=120 * H * W / sec, unit="rays/s", frames_per_s=120 / sec, seconds=sec, n_gpus=world,
         chunk=chunk, sharding="frame-parallel, round-robin")
if world > 1:
    dist.destroy_process_group()
