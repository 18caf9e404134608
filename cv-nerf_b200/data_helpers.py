"""Drop-in for the hot-path parts of the reference's ``data_helpers.py``: get_ndc and the
spherical pose generator used for the blender render path.  Dataset loaders (images, COLMAP
poses) are host I/O outside the accelerated path and are not reproduced here."""
import numpy as np
import torch

from . import kernels as K
from ._lib import NerfB200Error


def get_ndc(height, width, focal, near, r_ori, r_dir):
    """NDC re-parameterisation exactly as the reference computes it, quirks included
    (/root/reference/data_helpers.py:327-344).  Bit-exact against the reference's CPU result."""
    if not r_dir.is_cuda:
        raise NerfB200Error("get_ndc needs CUDA tensors; there is no CPU fallback")
    o = r_ori.expand(r_dir.shape) if r_ori.shape != r_dir.shape else r_ori
    return K.get_ndc(int(height), int(width), focal, near, o, r_dir)


def _translate_z(t):
    m = torch.eye(4)
    m[2, 3] = t
    return m


def _rot_phi(phi):
    c, s = np.cos(phi), np.sin(phi)
    return torch.tensor([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=torch.float32)


def _rot_theta(th):
    c, s = np.cos(th), np.sin(th)
    return torch.tensor([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=torch.float32)


def pose_spherical(theta, phi, radius):
    """Camera-to-world pose on a sphere (/root/reference/data_helpers.py:34-41); host-side 4x4."""
    c2w = _rot_theta(theta / 180. * np.pi) @ (_rot_phi(phi / 180. * np.pi) @ _translate_z(radius))
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ c2w
