"""Headline benchmark: rendered rays/s (64 coarse + 128 fine samples) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): lego 800x800 full-res view, f=1111.111, spherical poses
(radius 4, phi -30), near/far 2/6, white background, no NDC, random-init weights (torch seed 0),
synthetic data.  One step = one full frame: ray generation -> coarse depths -> coarse field ->
compositing -> inverse-CDF resampling + merge -> fine field -> compositing.  With N > 1 the frame's
rows are sharded over the ranks (rays are independent; strong scaling) and every rank's slice is
all-gathered so each step ends with the whole frame on every rank.

Printed JSON (rank 0): `value` = rays/s with inputs resident in HBM; `e2e` = the same through the
public render() call with the pose coming from pinned host memory and the frame read back to
pinned host memory inside the timed region; `roofline` = the field-network kernel against the
measured BF16 tensor peak; `cpu_baseline` = the reference's CPU path on a bounded sample of the same
frame; `train` = BASELINE.json configs[3] (4096-ray data-parallel train step) with its own HBM
roofline; `configs` = the fern (configs[2]) and skull-video (configs[4]) shapes.

`--impl reference` times the reference's own `main.render` on the host CPUs, all host threads, on
bounded samples of the same frame: the reference is pure Python, /root/reference does not exist on
the GPU box, so its modules are byte-compiled in the build container into oracle/_ref/
(oracle/build_ref.py; git-ignored, travels with the snapshot) -- kind "reference"; if that directory
is absent the oracle port (pinned to the reference's outputs by tests/) is timed -- kind "port".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 800
FOCAL = 1111.1110311937682      # .5 * 800 / tan(.5 * 0.6911112070083618), data_helpers.py:88
NEAR, FAR = 2., 6.
N_COARSE, N_FINE = 64, 128
FLOP_PER_RAY = (N_COARSE + N_COARSE + N_FINE) * 2 * 593408
WORKLOAD = "lego 800x800 full-res render, 64 coarse + 128 fine samples, white_bkg, random-init weights"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d, "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


def poses(n, pose_fn):
    """The blender render path: 40 poses on a circle, phi = -30, radius 4 (data_helpers.py:91)."""
    return [pose_fn(float(a), -30., 4.)[:3, :4].contiguous() for a in np.linspace(-180., 180., n + 1)[:-1]]


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                 "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so
    torch.get_num_threads() is 1 there; the CPU arm must not inherit that."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def cpu_reference_rays_per_s(n_rays, steps, warmup):
    """The reference's CPU path on `n_rays` rays of the benchmark frame per step, fp32, all host
    threads -> (rays/s, ms_per_step, threads, kind).  kind "reference": the reference's own main.render
    run from oracle/_ref (its modules byte-compiled in the build container by oracle/build_ref.py);
    kind "port": the oracle restatement (tests pin it to the reference's outputs) when oracle/_ref did
    not travel."""
    from oracle import build_ref
    from oracle import nerf_oracle as O
    threads = host_threads()
    torch.manual_seed(0)
    pose_list = poses(40, O.lego_pose)
    idx = torch.arange(n_rays) * ((H * W) // n_rays)             # bounded, strided sample of the frame
    if build_ref.available():
        from types import SimpleNamespace
        ref = build_ref.load()
        cfg = SimpleNamespace(netchunk=65536, lr=5e-4, perturb=0., n_fine_samples=N_FINE, n_coarse_samples=N_COARSE,
                              white_bkg=True, noise=0., dtype="blender", no_ndc=False)
        _, kw_test, _, _, _ = ref.create_model(cfg)              # torch.manual_seed(0) default init, main.py:127-167
        kw_test.update(near=NEAR, far=FAR)

        def one(pose):
            o, d = ref.compute_rays(H, W, FOCAL, pose)
            rays = torch.stack([o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]], 0)
            t0 = time.perf_counter()
            ref.render(H, W, FOCAL, chunk=32768, rays=rays, **kw_test)          # main.py:49-87
            return time.perf_counter() - t0
        kind = "reference"
    else:
        coarse, fine = O.init_field_params(0)

        def one(pose):
            o, d = O.ray_grid(H, W, FOCAL, pose)
            rays = (o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx])
            t0 = time.perf_counter()
            O.render_image(H, W, FOCAL, coarse, fine, rays=rays, ndc=False, near=NEAR, far=FAR,
                           n_coarse=N_COARSE, n_fine=N_FINE, white_bkg=True)
            return time.perf_counter() - t0
        kind = "port"
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            dt = one(pose_list[it % len(pose_list)])
            if it >= warmup:
                times.append(dt)
    total = sum(times)
    return n_rays * len(times) / total, 1e3 * total / len(times), threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = args.ref_rays
    rps, ms, threads, kind = cpu_reference_rays_per_s(n_rays, args.steps, max(args.warmup, 1))
    sample = f"{n_rays} rays per step, strided over the {H}x{W} frame, full 64+128 pipeline, fp32, {threads} threads"
    line = {
        "impl": "reference", "metric": "rendered rays/sec (64+128 samples)", "value": rps, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    import cv_nerf_b200
    from cv_nerf_b200 import main as M
    from cv_nerf_b200 import kernels as K
    from cv_nerf_b200.model import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a short watchdog: a rank that misses a collective must fail the run, not hang the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))

    torch.manual_seed(0)                      # same weights on every rank (create_model order)
    cfg = M.load_config(None, dtype="blender", white_bkg=True, n_coarse_samples=N_COARSE, n_fine_samples=N_FINE)
    _, kw_test, _, _, _ = M.create_model(cfg)
    kw_test.update(near=NEAR, far=FAR)
    from cv_nerf_b200.data_helpers import pose_spherical
    pose_host = torch.stack(poses(40, pose_spherical)).pin_memory()          # [40,3,4] pinned
    pose_dev = pose_host.to(dev)

    # contiguous row blocks per rank (rays are independent: no exchange inside the render)
    from cv_nerf_b200 import parallel as P
    bounds = P.row_bounds(H, world)                   # blocks differ by at most one row (ragged allowed)
    r0, r1 = bounds[rank], bounds[rank + 1]
    frame = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    host_out = torch.empty((r1 - r0, W, 3), dtype=torch.float32).pin_memory()
    pose_slot = torch.empty((3, 4), dtype=torch.float32, device=dev)

    def step_device(i):
        rgb, _ = M.render(H, W, FOCAL, c2w=pose_dev[i % 40], rows=(r0, r1), **kw_test)
        if world > 1:
            P.all_gather_rows(rgb, H)                 # the whole frame on every rank
        return rgb

    def step_e2e(i):
        pose_slot.copy_(pose_host[i % 40], non_blocking=True)              # H2D from pinned memory
        rgb, _ = M.render(H, W, FOCAL, c2w=pose_slot, rows=(r0, r1), **kw_test)
        host_out.copy_(rgb, non_blocking=True)                             # D2H of this rank's rows
        torch.cuda.current_stream().synchronize()                          # the caller holds the pixels
        return rgb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, timed_kernels):
        with torch.no_grad():
            for i in range(warmup):
                step_fn(i)
            barrier()
            K.STATS.reset(timed=timed_kernels)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step_fn(warmup + i)
            e1.record()
            barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local)
    with sampler:
        total_ms = timed(step_device, args.steps, args.warmup, timed_kernels=True)
    launches = K.STATS.launches
    kern = K.STATS.timed or []
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in kern)
    kern_rows = sum(r for _, _, r in kern)
    K.STATS.reset()
    e2e_ms = timed(step_e2e, args.steps, max(args.warmup, 1), timed_kernels=False)

    train = None
    if not args.no_train:
        train = bench_train_step(args, dev, world, rank)

    extra = None
    if not args.no_extra_configs:
        extra = bench_extra_configs(dev, world, rank)

    peaks, peak_kind = measured_peaks()
    n_rays = H * W
    rays_per_s = n_rays * args.steps / (total_ms * 1e-3)
    e2e_rays_per_s = n_rays * args.steps / (e2e_ms * 1e-3)
    achieved = kern_rows * K.FLOP_PER_SAMPLE / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.
    peak = float(peaks["bf16_tflops_sustained"])   # kernel timed inside a long step
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_rays
        rps, ms, threads, kind = cpu_reference_rays_per_s(n_cpu, 1, 1)
        cpu = {"value": rps, "unit": "rays/s", "cores": threads, "kind": kind,
               "sample": f"{n_cpu} rays of the same frame (strided), 1 warm-up + 1 timed pass, torch-CPU fp32"}
    line = {
        "metric": "rendered rays/sec (64+128 samples)", "value": rays_per_s, "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n_rays, "sharding": f"{world} row blocks of {min(b - a for a, b in zip(bounds[:-1], bounds[1:]))}..{max(b - a for a, b in zip(bounds[:-1], bounds[1:]))} image rows",
                   "random_draws": "in-kernel Philox (resampling uniforms; no torch.rand launch, no u tensor in HBM)",
                   "l2": "per-step intermediates (1.97 GB raw + 0.49 GB depths) exceed the 126 MB L2; no flush needed",
                   "weights": "torch.manual_seed(0) default nn.Linear init"},
        "e2e": {"value": e2e_rays_per_s, "unit": "rays/s", "h2d_bytes_per_step": 48 * world,
                "d2h_bytes_per_step": n_rays * 12, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak,
                     "frac_note": "can exceed 1: `achieved` counts the reference's multiply-adds (l9 included) while the kernel "
                                  "executes 0.89x of them (executed_tflops), and `peak` is the SUSTAINED cuBLAS figure of "
                                  "MEASURED_PEAKS.json, taken at the clock cuBLAS reaches under the 1 kW cap; this kernel keeps "
                                  "weights and activations on chip, draws less power per FLOP and holds a higher clock "
                                  "(clocks.sm_mhz). Against the burst peak: frac_of_burst_peak (algorithmic), "
                                  "executed_frac_of_burst_peak",
                     # DRAM bytes per launch: ncu --set full measured 135.3 MB read + 255.7 MB written for a
                     # 19.2 M-row launch of this kernel (profiles/r02_fwd_ncu.txt) = 20.4 B/row, against
                     # ~23 B/row algorithmic (depth in, raw out, rays/view term per ray): no re-reads
                     "traffic": NCU_DRAM_BYTES_PER_ROW * kern_rows / max(len(kern), 1),
                     "traffic_unit": "bytes per launch",
                     "traffic_source": "ncu --set full capture of this kernel (profiles/r02_fwd_ncu.txt: dram__bytes_read.sum "
                                       "+ dram__bytes_write.sum = 20.4 B per row) x rows per launch of this run; not re-measured inside bench.py",
                     "kernel": "mlp_fwd_kernel" if os.environ.get("NERF_B200_FWD_PAIRS", "1")[:1] == "0" else "mlp_fwd_pair_kernel",
                     # `achieved` counts the REFERENCE's multiply-adds (593 408 per sample).  The kernel
                     # executes fewer: l9 has no activation and is folded into l10 (csrc/mlp_layout.h), 527 872 per sample
                     "executed_tflops": achieved * EXECUTED_MAC_PER_SAMPLE / 593408.,
                     "executed_frac_of_peak": achieved * EXECUTED_MAC_PER_SAMPLE / 593408. / peak,
                     "peak_kind": f"{peak_kind} bf16_tflops_sustained",
                     "frac_of_burst_peak": achieved / float(peaks["bf16_tflops"]),
                     "executed_frac_of_burst_peak": achieved * EXECUTED_MAC_PER_SAMPLE / 593408. / float(peaks["bf16_tflops"]),
                     "kernel_share_of_step": kern_ms / total_ms,
                     "whole_step_tflops": rays_per_s * FLOP_PER_RAY / 1e12},
        "clocks": sampler.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if train is not None:
        line["train"] = train
    if extra is not None:
        line["configs"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


EXECUTED_MAC_PER_SAMPLE = 593408 - 65536 - 32768 + 32768     # l9 (256x256) folded into l10[:, :256] (128x256)
NCU_DRAM_BYTES_PER_ROW = (136.243e6 + 255.897e6) / 19.2e6     # profiles/r02_fwd_ncu.txt (mlp_fwd_pair_kernel)
TRAIN_FLOP_PER_RAY = (N_COARSE + N_COARSE + N_FINE) * 2 * (593408 + 593408 + 557696)


def bench_train_step(args, dev, world, rank):
    """BASELINE.json configs[3]: lego training, 4096 rays per GPU, data parallel: device-side batch,
    forward with saved activations, loss, backward, all-reduce of the gradient blobs, Adam, re-pack.
    Returns the `train` object of the JSON line (ms per step = max over ranks)."""
    import torch.distributed as dist
    from cv_nerf_b200 import kernels as K
    from cv_nerf_b200.data_helpers import pose_spherical
    from cv_nerf_b200.model import Model
    from cv_nerf_b200.train import TrainStep
    torch.manual_seed(0)
    coarse, fine = Model().to(dev), Model().to(dev)        # same init on every rank (seeded)
    h = w = 400
    focal = FOCAL / 2.
    n_rays = args.train_rays
    ts = TrainStep(coarse, fine, height=h, width=w, focal=focal, n_rays=n_rays, perturb=1., noise=0., white_bkg=True,
                   ndc=False, near=NEAR, far=FAR, lr=5e-4, lr_decay=500, seed=1234 + rank)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    images = torch.rand((4, h, w, 3), device=dev, generator=g)            # synthetic targets, per rank
    pose_dev = torch.stack(poses(4, pose_spherical)).to(dev)
    steps, warmup = args.train_steps, max(args.warmup, 3)

    def one(i):
        return ts.step(images[i % 4], pose_dev[i % 4])

    for i in range(warmup):
        one(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    K.STATS.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = one(warmup + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = K.STATS.launches
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / steps
    # where the data-parallel step spends its time: forward+backward vs gradient exchange + Adam + re-pack
    # (CUDA events inside TrainStep.step over 10 more steps, max over ranks; not part of the timed region)
    ts.profile = []
    for i in range(10):
        one(warmup + steps + i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    halves = torch.tensor([sum(e[0].elapsed_time(e[1]) for e in ts.profile) / len(ts.profile),
                           sum(e[1].elapsed_time(e[2]) for e in ts.profile) / len(ts.profile)], device=dev)
    ts.profile = None
    if world > 1:
        dist.all_reduce(halves, op=dist.ReduceOp.MAX)
    fwd_bwd_ms, update_ms = (float(v) for v in halves.tolist())
    # per-kernel split of one more step (events around each stage; not part of the timed region)
    stages = {}
    if rank == 0:
        stages = time_train_stages(ts, images[0], pose_dev[0])
    peaks, _ = measured_peaks()
    tflops = n_rays * TRAIN_FLOP_PER_RAY / (ms_step * 1e-3) / 1e12
    # HBM roofline of the step: the training kernels are HBM-bound by design --
    # the forward writes the activation records, the dZ chain writes the dZ records, the dW kernel reads
    # both.  Algorithmic bytes = every record written once and read once (+ raw / grad_raw, 32 B/row).
    rows = n_rays * (N_COARSE + N_COARSE + N_FINE)
    rec_bytes = sum(K.act_bytes(n_rays * s_) + K.dz_bytes(n_rays * s_) for s_ in (N_COARSE, N_COARSE + N_FINE))
    hbm_bytes = 2 * rec_bytes + 32 * rows
    hbm_gbs = hbm_bytes / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": hbm_gbs, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                "frac": hbm_gbs / float(peaks["hbm_gbs"]), "traffic": None,
                "bytes_per_step": hbm_bytes, "bytes_per_row": hbm_bytes / rows,
                "what": "activation + dZ tile records written once and read once per step, whole step time "
                        "(forward, compositing, loss, dZ chains, dW, Adam, re-pack); write-only HBM measures "
                        "3.9 TB/s on this part, so a step that writes half of its bytes cannot reach 1.0",
                "tensor_frac_of_sustained_bf16_peak": tflops / float(peaks["bf16_tflops_sustained"])}
    return {"roofline": roofline,"metric": "train step ms (4096 rays per GPU, fwd+bwd+allreduce+Adam)", "ms_per_step": ms_step,
            "rays_per_step_per_gpu": n_rays, "rays_per_s_all_gpus": n_rays * world / (ms_step * 1e-3),
            "steps": steps, "warmup": warmup, "tflops_per_gpu": tflops,
            "frac_of_sustained_bf16_peak": tflops / float(peaks["bf16_tflops_sustained"]),
            "gpu_launches_per_step": launches / steps, "loss_last": float(loss.item()),
            "forward_backward_ms": fwd_bwd_ms, "exchange_adam_repack_ms": update_ms,
            "grad_exchange_bytes_per_rank": int(ts.blob.numel() * 4) if world > 1 else 0,
            "grad_exchange": ("none" if world == 1 else
                              "fused into Adam: peer loads over NVLink from symmetric memory" if ts.symm is not None
                              else "nccl all_reduce"),
            "stages_ms": stages}


def bench_extra_configs(dev, world, rank):
    """BASELINE.json configs[2] and configs[4] on the same GPUs (max over ranks, CUDA events):
      fern 378x504 NDC: one frame (row-sharded like the headline) and a train step with the
      configs/fern.txt flags (perturb 1, noise 1);
      skull 504x378 NDC: the reference's 120 spiral poses (tests/golden/skull_spiral.npz, derived with
      its own pose math) through render_full, frame-parallel over the ranks, uint8 frames to the host.
    Fern poses are not in the reference tree: the fern shape is rendered from a
    recentred skull training pose."""
    import torch.distributed as dist
    from cv_nerf_b200 import main as M
    from cv_nerf_b200 import parallel as P
    from cv_nerf_b200.model import Model
    from cv_nerf_b200.train import TrainStep
    path = os.path.join(ROOT, "tests", "golden", "skull_spiral.npz")
    if not os.path.exists(path):
        return {"unavailable": "tests/golden/skull_spiral.npz missing"}
    g = np.load(path)
    torch.manual_seed(0)
    coarse, fine = Model().to(dev), Model().to(dev)
    kw = dict(coarse_model=coarse, fine_model=fine, n_coarse_samples=N_COARSE, n_fine_samples=N_FINE, white_bkg=False,
              ndc=True, near=0., far=1., perturb=False, noise=0.)

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / reps

    out = {}
    h, w, f = 378, 504, np.float32(407.5657)
    pose = torch.from_numpy(g["train_poses"][3][:3, :4]).float().to(dev)
    b = P.row_bounds(h, world)

    def fern_frame():
        with torch.no_grad():
            rgb, _ = M.render(h, w, f, c2w=pose, rows=(b[rank], b[rank + 1]), **kw)
            P.all_gather_rows(rgb, h)
    ms = timed(fern_frame, 10)
    out["fern_render"] = {"workload": "fern 378x504 NDC frame, 64+128 samples, rows sharded over the GPUs",
                          "value": h * w / (ms * 1e-3), "unit": "rays/s", "ms_per_frame": ms, "n_gpus": world}
    ts = TrainStep(coarse, fine, height=h, width=w, focal=f, n_rays=4096, perturb=1., noise=1., white_bkg=False, ndc=True,
                   near=0., far=1., lr=5e-4, lr_decay=250, seed=77)
    image = torch.rand(h, w, 3, device=dev)
    ms = timed(lambda: ts.step(image, pose), 20, warm=3)
    out["fern_train"] = {"workload": "fern train step (configs/fern.txt: perturb 1, noise 1), 4096 rays per GPU, data parallel",
                         "ms_per_step": ms, "rays_per_s_all_gpus": 4096 * world / (ms * 1e-3), "n_gpus": world}
    del ts
    h, w, f = int(g["hwf"][0]), int(g["hwf"][1]), np.float32(g["hwf"][2])
    spiral = [torch.from_numpy(p_).float().to(dev) for p_ in g["render_poses"]]
    kw_v = dict(kw)
    sec = timed(lambda: M.render_full(spiral, [h, w, f], 32768, kw_v, as_bytes=True, verbose=False), 1, warm=0) * 1e-3
    out["skull_video"] = {"workload": "skull 504x378 NDC, the reference's 120-frame spiral, frame-parallel over the GPUs, "
                                      "uint8 frames to pinned host memory (D2H inside the timed region)",
                          "value": len(spiral) * h * w / sec, "unit": "rays/s", "frames_per_s": len(spiral) / sec,
                          "seconds": sec, "n_gpus": world}
    return out


def time_train_stages(ts, image, pose):
    """CUDA-event time of each stage of one train step (diagnostic split, same kernels)."""
    from cv_nerf_b200 import kernels as K
    ev = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append((name, e))
    n, dev = ts.n_rays, ts.dev
    mark("start")
    rays, target, _ = K.train_rays(ts.h, ts.w, ts.f, pose, n, seed=1, image=image, ndc=ts.ndc, near=ts.near, far=ts.far)
    z_c = K.sample_coarse(rays, ts.s_c, torch.rand((n, ts.s_c), device=dev))
    u = torch.rand((n, ts.n_fine), device=dev)
    pk_c, pk_f = ts.coarse.packed(), ts.fine.packed()
    vt_c, vt_f = K.viewdir_term(pk_c, rays), K.viewdir_term(pk_f, rays)
    mark("batch+sampling")
    raw_c = K.mlp_fwd(pk_c, K.IN_RAYS, rays, z_c, n * ts.s_c, ts.s_c, vt_c, ts.s_c, act_save=ts.act_c)
    mark("fwd_coarse(save)")
    rgb_c, w_c = K.composite_fwd(raw_c.view(n, ts.s_c, 4), z_c, rays, None, ts.white_bkg)
    z_f = K.resample_merge(z_c, w_c, u)
    mark("composite+resample")
    raw_f = K.mlp_fwd(pk_f, K.IN_RAYS, rays, z_f, n * ts.s_f, ts.s_f, vt_f, ts.s_f, act_save=ts.act_f)
    mark("fwd_fine(save)")
    rgb_f, _ = K.composite_fwd(raw_f.view(n, ts.s_f, 4), z_f, rays, None, ts.white_bkg, want_weights=False)
    ts.loss.zero_()
    _, g_f = K.mse_loss_grad(rgb_f, target, loss=ts.loss)
    _, g_c = K.mse_loss_grad(rgb_c, target, loss=ts.loss)
    graw_f = K.composite_bwd(raw_f.view(n, ts.s_f, 4), z_f, rays, None, ts.white_bkg, g_f).view(-1, 4)
    graw_c = K.composite_bwd(raw_c.view(n, ts.s_c, 4), z_c, rays, None, ts.white_bkg, g_c).view(-1, 4)
    ts.blob.zero_()
    mark("composite+loss+bwd")
    K.mlp_bwd_dz(ts.fine.packed_bwd(), graw_f, ts.act_f, n * ts.s_f, dz=ts.dz)
    mark("dz_fine")
    K.mlp_bwd_params(ts.act_f, ts.dz, graw_f, n * ts.s_f, rays, ts.s_f, False, ts.blob[1], params=ts.params[1])
    mark("dw_fine(+heads,view)")
    K.mlp_bwd_dz(ts.coarse.packed_bwd(), graw_c, ts.act_c, n * ts.s_c, dz=ts.dz)
    mark("dz_coarse")
    K.mlp_bwd_params(ts.act_c, ts.dz, graw_c, n * ts.s_c, rays, ts.s_c, False, ts.blob[0], params=ts.params[0])
    mark("dw_coarse(+heads,view)")
    ts.apply_gradients(allreduce=False)     # this split runs on rank 0 only: no collectives here
    ts.coarse.packed(); ts.fine.packed(); ts.coarse.packed_bwd(); ts.fine.packed_bwd()
    mark("adam+repack")
    torch.cuda.synchronize()
    return {name: ev[i - 1][1].elapsed_time(e) for i, (name, e) in enumerate(ev) if i > 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=8192, help="rays of the CPU-baseline sample (ours arm)")
    ap.add_argument("--ref-rays", type=int, default=2048, help="rays per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step measurement")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the fern / skull-video measurements")
    ap.add_argument("--train-rays", type=int, default=4096, help="rays per GPU and train step (BASELINE configs[3])")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--size", type=int, default=800, help="frame height=width (profiling runs only; 800 is the benchmark)")
    args = ap.parse_args()
    global H, W, FOCAL, WORKLOAD
    if args.size != 800:
        H = W = args.size
        FOCAL = FOCAL * args.size / 800.
        WORKLOAD = WORKLOAD.replace("800x800 full-res", f"{H}x{W} (NOT the benchmark size)")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
