"""Multi-GPU paths on real devices (needs >= 2 visible GPUs; skipped on the 1-GPU box): one process
per GPU over NCCL on 127.0.0.1.

* a frame whose rows are sharded over 2 ranks equals the 1-rank frame bit for bit (main.py:49-87;
  rays are independent, the in-kernel draws are keyed by the global ray index);
* render_full's frame-parallel stack equals the single-rank stack (main.py:102-124);
* data-parallel TrainStep: ranks draw different batches, the fused peer-memory exchange+Adam path and
  the NCCL all-reduce path give the same parameters (to the floating-point-atomics floor), replicas stay
  bit-identical (main.py:376-394).
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import datetime
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=120))
    try:
        import cv_nerf_b200  # noqa: F401
        from cv_nerf_b200 import kernels as K, main as M, parallel as P
        from cv_nerf_b200.data_helpers import pose_spherical
        from cv_nerf_b200.model import Model
        from cv_nerf_b200.train import TrainStep
        out = {}
        torch.manual_seed(0)
        coarse, fine = Model().to(dev), Model().to(dev)
        with torch.no_grad():
            coarse.l_alpha.bias.fill_(1.); coarse.l_alpha.weight.mul_(5.)
            fine.l_alpha.bias.fill_(1.); fine.l_alpha.weight.mul_(5.)
        kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=64, n_fine_samples=128, white_bkg=True,
                  ndc=False, near=2., far=6., perturb=1., noise=0.25)

        # 1. row-sharded frame (ragged split: 37 rows over 2 ranks) == whole frame
        h, w, f = 37, 48, 60.
        pose = pose_spherical(30., -30., 4.)[:3, :4].to(dev)
        rng = K.Rng(1234)
        b = P.row_bounds(h, world)
        with torch.no_grad():
            mine, _ = M.render(h, w, f, c2w=pose, rows=(b[rank], b[rank + 1]), rng=rng, **kw)
            gathered = P.all_gather_rows(mine, h)
            whole, _ = M.render(h, w, f, c2w=pose, rng=rng, **kw)
        out["rows_equal"] = bool(torch.equal(gathered, whole))
        out["frame_mean"] = float(whole.mean())

        # 2. frame-parallel video == single-rank video (5 frames over 2 ranks; same seed on both ranks)
        poses = [pose_spherical(a, -30., 4.).to(dev) for a in (-180., -100., -20., 60., 140.)]
        kw_test = dict(kw, perturb=False, noise=0.)

        torch.manual_seed(5)
        vid = M.render_full(poses, [20, 24, 30.], 32768, kw_test, as_bytes=True, verbose=False)
        out["video_shape"] = tuple(vid.shape)
        # every rank returns the full stack, and the stacks agree across ranks
        mine_v = torch.from_numpy(vid).to(dev)
        both = [torch.empty_like(mine_v) for _ in range(world)]
        dist.all_gather(both, mine_v)
        out["video_same_on_all_ranks"] = bool(torch.equal(both[0], both[1]))
        # frame i was rendered by rank i % world as that rank's (i // world)-th render() call
        torch.manual_seed(5)
        my_frames = P.frame_indices(len(poses), world, rank)
        ok = True
        with torch.no_grad():
            for i in my_frames:
                frame = K.to_byte(M.render(20, 24, 30., c2w=poses[i][:3, :4], **kw_test)[0])
                ok &= bool(torch.equal(frame, mine_v[i]))
        out["video_frames_equal_single_rank_renders"] = ok

        # 3. data-parallel training
        def run(peer):
            os.environ["NERF_B200_PEER_ADAM"] = "1" if peer else "0"
            torch.manual_seed(0)
            c, fi = Model().to(dev), Model().to(dev)
            ts = TrainStep(c, fi, height=64, width=64, focal=90., n_rays=512, perturb=1., noise=0., white_bkg=True,
                           ndc=False, near=2., far=6., seed=0)          # default-style seed: the rank is folded in
            g = torch.Generator(device=dev).manual_seed(7 + rank)
            image = torch.rand(64, 64, 3, device=dev, generator=g)
            p = pose_spherical(-180. + 20 * rank, -30., 4.)[:3, :4].to(dev)
            losses = [ts.step(image, p).item() for _ in range(3)]
            params = torch.cat([q.detach().reshape(-1) for q in list(c.parameters()) + list(fi.parameters())])
            return ts, params, losses
        ts_p, p_peer, l_peer = run(True)
        ts_n, p_nccl, l_nccl = run(False)
        _, p_nccl2, _ = run(False)
        seeds_t = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(seeds_t, torch.tensor([ts_p.seed], dtype=torch.int64, device=dev))
        out["rank_seeds_differ"] = len({int(t) for t in seeds_t}) == world
        _, _, pix = K.train_rays(64, 64, 90., pose, 512, seed=ts_p.seed * 0x9E3779B97F4A7C15, want_pix=True, ndc=False)
        pix_all = [torch.empty_like(pix) for _ in range(world)]
        dist.all_gather(pix_all, pix)
        out["rank_batches_differ"] = not torch.equal(pix_all[0], pix_all[1])
        out["peer_path_active"] = ts_p.symm is not None
        out["peer_vs_nccl"] = float((p_peer - p_nccl).abs().max())
        out["nccl_run_to_run"] = float((p_nccl - p_nccl2).abs().max())
        out["losses_close"] = all(abs(a - c_) <= 1e-5 * max(1., abs(c_)) for a, c_ in zip(l_peer, l_nccl))
        reps = [torch.empty_like(p_peer) for _ in range(world)]
        dist.all_gather(reps, p_peer)
        out["replicas_identical"] = bool(torch.equal(reps[0], reps[1]))
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_two_rank_render_video_and_data_parallel_training():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    res = dict(ret)
    print(res)
    from tests.helpers import record
    record("multirank", {f"rank{r}": v for r, v in res.items()})
    for r in (0, 1):
        o = res[r]
        assert o["rows_equal"], "row-sharded frame differs from the whole frame"
        assert 0.05 < o["frame_mean"] < 0.95
        assert o["video_shape"] == (5, 20, 24, 3) and o["video_same_on_all_ranks"]
        assert o["video_frames_equal_single_rank_renders"]
        assert o["rank_seeds_differ"] and o["rank_batches_differ"]
        assert o["losses_close"] and o["replicas_identical"]
        assert o["peer_vs_nccl"] <= max(4 * o["nccl_run_to_run"], 1.5e-3), o
