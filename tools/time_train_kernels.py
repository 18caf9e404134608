"""CUDA-event timing of every kernel of one train iteration, each launched alone `reps` times
(BASELINE.json configs[3] shape).  Usage: python tools/time_train_kernels.py [n_rays] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200 import _lib, kernels as K  # noqa: E402
from cv_nerf_b200.data_helpers import pose_spherical  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402
from cv_nerf_b200.train import TrainStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
coarse, fine = Model().to(dev), Model().to(dev)
ts = TrainStep(coarse, fine, height=400, width=400, focal=555.5555, n_rays=n, perturb=1., noise=0., white_bkg=True,
               ndc=False, near=2., far=6., seed=1)
image = torch.rand(400, 400, 3, device=dev)
pose = pose_spherical(-180., -30., 4.)[:3, :4].to(dev)
for _ in range(2):
    ts.step(image, pose)
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def timeit(name, fn, bytes_=None, flop=None):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = ""
    if bytes_:
        extra += f"  {bytes_ / ms / 1e9:.2f} TB/s"
    if flop:
        extra += f"  {flop / ms / 1e9:.0f} TFLOP/s"
    print(f"{name:28s} {ms:8.3f} ms{extra}")
    return ms


rays, target, _ = K.train_rays(400, 400, 555.5555, pose, n, seed=1, image=image, ndc=False, near=2., far=6.)
total = 0.
for tag, net, S, act in (("coarse", coarse, 64, ts.act_c), ("fine", fine, 192, ts.act_f)):
    rows = n * S
    z = K.sample_coarse(rays, S, torch.rand(n, S, device=dev))
    pk, pkb = net.packed(), net.packed_bwd()
    vt = K.viewdir_term(pk, rays)
    graw = torch.randn(rows, 4, device=dev) * 1e-3
    blob = torch.zeros(K.grad_blob_floats(), device=dev)
    total += timeit(f"fwd {tag}", lambda: K.mlp_fwd(pk, K.IN_RAYS, rays, z, rows, S, vt, S), flop=rows * 1186816)
    total += timeit(f"fwd+save {tag}", lambda: K.mlp_fwd(pk, K.IN_RAYS, rays, z, rows, S, vt, S, act_save=act),
                    bytes_=rows * 4992, flop=rows * 1186816) * 0
    total += timeit(f"dz {tag}", lambda: K.mlp_bwd_dz(pkb, graw, act, rows, dz=ts.dz), bytes_=rows * (288 + 4864),
                    flop=rows * 2 * 557696)
    total += timeit(f"dw {tag}", lambda: _lib.check(lib.nerf_mlp_bwd_dw(act.data_ptr(), ts.dz.data_ptr(), rows, blob.data_ptr(), st), "dw"),
                    bytes_=rows * 9856, flop=rows * 2 * 592768)
    timeit(f"heads {tag}", lambda: _lib.check(lib.nerf_mlp_bwd_heads(act.data_ptr(), graw.data_ptr(), rows, blob.data_ptr(), st), "heads"),
           bytes_=rows * 768)
    timeit(f"viewdir_bwd {tag}", lambda: _lib.check(lib.nerf_viewdir_term_bwd(ts.dz.data_ptr(), rays.data_ptr() + 32, 11, 0, rows, S, blob.data_ptr(), st), "v"),
           bytes_=rows * 256)
timeit("pack_model", lambda: K.pack_model(fine.ordered_params()))
timeit("pack_model_bwd", lambda: K.pack_model_bwd(fine.ordered_params()))
timeit("viewdir_term", lambda: K.viewdir_term(fine.packed(), rays))
timeit("train_rays", lambda: K.train_rays(400, 400, 555.5555, pose, n, seed=1, image=image, ndc=False, near=2., far=6.))
timeit("blob.zero_", lambda: ts.blob.zero_())
timeit("adam_blob x2", lambda: [K.adam_step_blob(ts.blob[i], [p.data for p in ts.params[i]], ts.m[i], ts.v[i], 5e-4, (0.9, 0.999), 1e-8, 3) for i in range(2)])
timeit("full step", lambda: ts.step(image, pose))
