"""Two launches of the production inference field kernel (100000 rays x 192 samples) for an ncu capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cv_nerf_b200  # noqa: E402,F401
from cv_nerf_b200 import kernels as K  # noqa: E402
from cv_nerf_b200.model import Model  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Model().to(dev)
n, S = 100000, 192
rays = torch.zeros(n, 11, device=dev)
rays[:, 0:3] = torch.randn(n, 3, device=dev) * .3
rays[:, 3:6] = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
rays[:, 6], rays[:, 7] = 2., 6.
rays[:, 8:11] = rays[:, 3:6]
z = K.sample_coarse(rays, S)
pk = net.packed()
ht = net.host_tail()
vt = K.viewdir_term(pk, rays)
for _ in range(2):
    raw = K.mlp_fwd(pk, K.IN_RAYS, rays, z, n * S, S, vt, S, host_tail=ht)
torch.cuda.synchronize()
print("ok", float(raw[0, 0]))
