"""Drop-in for the reference's ``utils.py``: inv_transform_sampling, cont_to_byte8_im."""
import numpy as np
import torch

from . import kernels as K
from ._lib import NerfB200Error


def inv_transform_sampling(pts, weights, n, *, u=None):
    """Inverse-transform sampling of ``n`` points from the piecewise-constant pdf ``weights``
    over ``pts`` (/root/reference/utils.py:4-53).  pts [N,B], weights [N,B-1] -> [N,n], unsorted.
    The reference always draws ``torch.rand``; pass ``u`` to supply the uniforms."""
    if not pts.is_cuda:
        raise NerfB200Error("inv_transform_sampling needs CUDA tensors; there is no CPU fallback")
    if u is None:
        u = torch.rand(list(weights.shape[:-1]) + [n], device=pts.device)
    return K.sample_pdf(pts, weights.detach(), u.to(pts.device))


def cont_to_byte8_im(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)
