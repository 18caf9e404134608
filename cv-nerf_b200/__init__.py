"""B200-native drop-in for the NeRF ray-render hot path of johnfay11/CV-Nerf.

Sub-modules mirror the reference's module names so that
``from cv_nerf_b200.main import render, create_model`` reads like ``from main import ...``:

    main          compute_rays, render, batch_rays/batchify_rays, render_rays,
                  process_volume_info, create_model, render_full, load_config
    model         FreqEmbedding, Model, net_forward, combine, to_byte
    utils         inv_transform_sampling, cont_to_byte8_im
    data_helpers  get_ndc, pose_spherical

Every numerical function runs hand-written sm_100a CUDA from libnerf_b200.so through the C ABI
in include/nerf_b200.h.  There is no CPU implementation in this package.
"""
from . import _lib, kernels  # noqa: F401
from ._lib import NerfB200Error  # noqa: F401

__all__ = ["NerfB200Error", "kernels"]
