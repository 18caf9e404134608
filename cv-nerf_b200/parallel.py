"""Multi-GPU plumbing for the two ways the hot path shards: one process per
GPU, ``torch.distributed`` (NCCL on the box, gloo in the CPU tests) for the few bytes that move.

* Rendering: rays are independent, so a frame is cut into contiguous row blocks (each rank
  generates its own rays from the 12-float pose -- no input scatter) and the finished rows are
  all-gathered; a video is cut by frames with no communication at all until the final gather.
* Training: data parallel over the ray batch; each rank draws its own batch, the two flat gradient
  blobs are summed with one all-reduce and Adam divides by the world size (train.TrainStep).

Nothing here computes pixels; the functions work on whatever device the tensors live on.
"""
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def row_bounds(height, world):
    """Row block boundaries: rank r renders rows [b[r], b[r+1]).  Blocks differ by at most one row."""
    return [height * r // world for r in range(world + 1)]


def frame_indices(n_frames, world, rank):
    """Frames of a video rendered by `rank` (round-robin, so early frames finish first everywhere)."""
    return list(range(rank, n_frames, world))


def all_gather_rows(local_rows, height, group=None):
    """[rows_r, W, C] per rank -> the whole [height, W, C] frame on every rank."""
    rank, world = world_info(group)
    if world == 1:
        return local_rows
    b = row_bounds(height, world)
    w, c = local_rows.shape[1], local_rows.shape[2]
    if all(b[i + 1] - b[i] == b[1] - b[0] for i in range(world)):
        out = torch.empty((height, w, c), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out.view(-1), local_rows.reshape(-1).contiguous(), group=group)
        return out
    biggest = max(b[i + 1] - b[i] for i in range(world))
    pad = torch.zeros((biggest, w, c), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[i][:b[i + 1] - b[i]] for i in range(world)], 0)


def gather_frames(local_frames, n_frames, group=None):
    """Frames rendered round-robin (frame_indices) -> [n_frames, H, W, C] on every rank."""
    rank, world = world_info(group)
    if world == 1:
        return local_frames
    per = (n_frames + world - 1) // world
    shape = (per,) + tuple(local_frames.shape[1:])
    pad = torch.zeros(shape, dtype=local_frames.dtype, device=local_frames.device)
    pad[:local_frames.shape[0]] = local_frames
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((n_frames,) + tuple(local_frames.shape[1:]), dtype=local_frames.dtype, device=local_frames.device)
    for r in range(world):
        idx = frame_indices(n_frames, world, r)
        out[idx] = parts[r][:len(idx)]
    return out


def allreduce_sum_(flat, group=None):
    """In-place sum of a flat buffer across ranks (the gradient blobs of train.TrainStep)."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_parameters(modules, src=0, group=None):
    """Same initial weights everywhere (the reference has a single process; seeds are per rank)."""
    _, world = world_info(group)
    if world == 1:
        return
    for m in modules:
        for p in m.parameters():
            dist.broadcast(p.data, src=src, group=group)


def render_full_sharded(render_fn, render_poses, group=None):
    """Frame-parallel video render: ``render_fn(pose) -> [H,W,3]`` is called for this rank's frames
    only; the result is the full stack on every rank."""
    rank, world = world_info(group)
    mine = frame_indices(len(render_poses), world, rank)
    frames = [render_fn(render_poses[i]) for i in mine]
    local = torch.stack(frames, 0) if frames else None
    if local is None:   # more ranks than frames: contribute an empty stack of the right shape
        probe = render_fn(render_poses[0])
        local = probe.new_zeros((0,) + tuple(probe.shape))
    return gather_frames(local, len(render_poses), group)
