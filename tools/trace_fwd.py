"""Event trace of the production field kernel (probe build): time stamps of CTA 0's producer, MMA issuer and
the first warp of both epilogue groups over its 4th and 5th tile pair, printed as one table per sub-tile
layer.  Usage: python tools/trace_fwd.py [variant]   (1 = production layout, 11 = no weight streaming ...)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200 import _lib, kernels as K  # noqa: E402
from tools.gpu_diag import packed_model, DEV  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _lib.load()
if not _lib.has_experiments():
    sys.exit("needs the experiments build: make -C cv-nerf_b200/csrc experiments && "
             "NERF_B200_LIB=$PWD/cv-nerf_b200/libnerf_b200_exp.so python " + sys.argv[0])
p, packed = packed_model()
n_rays, S = 160000, 192
rays = torch.zeros(n_rays, 11, device=DEV)
rays[:, 0:3] = torch.randn(n_rays, 3, device=DEV) * .3
rays[:, 3:6] = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV), dim=-1)
rays[:, 6], rays[:, 7] = 2., 6.
rays[:, 8:11] = rays[:, 3:6]
z = K.sample_coarse(rays, S)
vt = K.viewdir_term(packed, rays)
raw = torch.empty(n_rays * S, 4, device=DEV)
buf = torch.zeros(148 * 8 + 4 * 1024, dtype=torch.int64, device=DEV)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    buf.zero_()
    rc = lib.nerf_mlp_fwd_stats(packed.data_ptr(), rays.data_ptr(), z.data_ptr(), n_rays * S, S, vt.data_ptr(),
                                raw.data_ptr(), 1000 + variant, buf.data_ptr(), st)
    torch.cuda.synchronize()
    assert rc == 0, lib.nerf_b200_last_error()
tr = buf[148 * 8:].view(4, 1024).cpu().tolist()
ev = []
for role in range(4):
    for e in tr[role]:
        if e == 0:
            continue
        ev.append((e & 0xFFFFFFFFFF, role, (e >> 56) & 0xff, (e >> 48) & 0xff, (e >> 44) & 0xf, (e >> 40) & 0xf))
ev.sort()
t0 = ev[0][0]
names = {(0, 1): "prod wait-free", (0, 2): "prod copy issued", (1, 1): "mma wait-A", (1, 2): "mma A ready", (1, 3): "mma wait-W",
         (1, 4): "mma W full", (1, 5): "mma chunk issued", (2, 1): "epiX wait-acc", (2, 2): "epiX acc done", (2, 3): "epiX A written",
         (3, 1): "epiY wait-acc", (3, 2): "epiY acc done", (3, 3): "epiY A written"}
if "--raw" in sys.argv:
    for t, role, tag, l, g, j in ev:
        print(f"{t - t0:8d}  {names[(role, tag)]:18s} l={l} g={g} j={j}")
# per sub-tile layer summary
print("layer g | A-wait  A-ready | per chunk: W-wait->W-full->issued ... | acc done (epi)  A written (epi) | span")
key = lambda l, g: [(t - t0, role, tag, j) for t, role, tag, ll, gg, j in ev if ll == l and gg == g]
seen = []
for t, role, tag, l, g, j in ev:
    if role == 1 and tag == 1 and (t, l, g) not in seen:
        seen.append((t, l, g))
for t_start, l, g in seen:
    rows = [(t - t0, role, tag, j) for t, role, tag, ll, gg, j in ev if ll == l and gg == g and t >= t_start]
    nxt = [t for t, ll, gg in seen if ll == l and gg == g and t > t_start]
    if nxt:
        rows = [r for r in rows if r[0] < nxt[0] - t0]
    a_wait = t_start - t0
    a_ready = next((t for t, role, tag, j in rows if role == 1 and tag == 2), None)
    chunks = []
    for jj in range(5):
        w0 = next((t for t, role, tag, j in rows if role == 1 and tag == 3 and j == jj), None)
        w1 = next((t for t, role, tag, j in rows if role == 1 and tag == 4 and j == jj), None)
        w2 = next((t for t, role, tag, j in rows if role == 1 and tag == 5 and j == jj), None)
        if w0 is not None:
            chunks.append(f"{w0 - a_wait}>{w1 - a_wait}>{w2 - a_wait}")
    acc = next((t for t, role, tag, j in rows if role == 2 + g and tag == 2), None)
    done = next((t for t, role, tag, j in rows if role == 2 + g and tag == 3), None)
    print(f"{l:2d} {g} | {a_wait:8d} +{(a_ready or a_wait) - a_wait:5d} | {'  '.join(chunks):70s} | "
          f"+{(acc or a_wait) - a_wait:6d}  +{(done or a_wait) - a_wait:6d}")
