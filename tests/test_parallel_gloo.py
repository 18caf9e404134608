"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: row/frame sharding arithmetic, the frame
gather, and the gradient all-reduce + divide that data-parallel training relies on."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cv_nerf_b200  # noqa: F401
        from cv_nerf_b200 import parallel as P
        ret[rank] = fn(rank, world, P)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _rows_case(rank, world, P):
    out = {}
    for height in (8, 7):           # even and ragged split
        full = torch.arange(height * 3 * 2, dtype=torch.float32).reshape(height, 3, 2)
        b = P.row_bounds(height, world)
        got = P.all_gather_rows(full[b[rank]:b[rank + 1]].clone(), height)
        out[height] = torch.equal(got, full)
    return out


def _frames_case(rank, world, P):
    poses = [torch.full((3, 4), float(i)) for i in range(5)]
    calls = []

    def render(pose):
        calls.append(int(pose[0, 0]))
        return pose[0, 0] * torch.ones(2, 2, 3)
    vid = P.render_full_sharded(render, poses)
    ok = all(torch.all(vid[i] == i).item() for i in range(5))
    return ok, calls


def _grad_case(rank, world, P):
    torch.manual_seed(100 + rank)
    blob = torch.randn(2, 1000)
    mine = blob.clone()
    P.allreduce_sum_(blob)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = sum(gathered)
    lin = torch.nn.Linear(4, 4)
    P.broadcast_parameters([lin])
    ws = [torch.empty_like(lin.weight.data) for _ in range(world)]
    dist.all_gather(ws, lin.weight.data)
    return torch.allclose(blob, want), bool(torch.equal(ws[0], ws[1]))


def test_row_sharded_frame_gather():
    res = _run(_rows_case)
    assert all(all(v.values()) for v in res.values()), res


def test_frame_parallel_video():
    res = _run(_frames_case)
    assert all(ok for ok, _ in res.values())
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]


def test_gradient_allreduce_and_broadcast():
    res = _run(_grad_case)
    assert all(a and b for a, b in res.values()), res


def test_sharding_arithmetic():
    from cv_nerf_b200 import parallel as P
    for h in (800, 378, 7):
        for world in (1, 2, 4, 8):
            b = P.row_bounds(h, world)
            assert b[0] == 0 and b[-1] == h and all(y >= x for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
    assert sorted(sum((P.frame_indices(120, 8, r) for r in range(8)), [])) == list(range(120))


def _seed_case(rank, world, P):
    from cv_nerf_b200.train import rank_seed
    mine = torch.tensor([rank_seed(0, rank), rank_seed(5, rank)], dtype=torch.int64)
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    return [t.tolist() for t in got]


def test_data_parallel_ranks_get_distinct_random_streams():
    """TrainStep folds the rank into the key of its pixel permutation and of its draw generator: with
    the default seed every rank would otherwise render the same batch (N GPUs doing the work of one)."""
    res = _run(_seed_case)
    seeds = res[0]
    assert seeds == res[1], "all_gather disagreement"
    assert seeds[0][0] != seeds[1][0] and seeds[0][1] != seeds[1][1]
    assert len({s for per_rank in seeds for s in per_rank}) == 4
