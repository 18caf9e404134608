"""ctypes binding of the C ABI declared in include/nerf_b200.h.

There is no CPU path: if libnerf_b200.so is missing or a tensor is not on a CUDA device the
call raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C cv-nerf_b200/csrc``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NERF_B200_LIB") or os.path.join(_HERE, "libnerf_b200.so")

c_float_p = ctypes.c_void_p  # device pointers travel as integers
_lib = None

# name -> (restype, argtypes); mirrors include/nerf_b200.h one to one
_SIGNATURES = {
    "nerf_b200_abi_version": (ctypes.c_int, []),
    "nerf_b200_last_error": (ctypes.c_char_p, []),
    "nerf_b200_sm_count": (ctypes.c_int, []),
    "nerf_compute_rays": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_float, c_float_p, ctypes.c_int,
                                         ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    "nerf_get_ndc": (ctypes.c_int, [ctypes.c_float, ctypes.c_float, ctypes.c_float, c_float_p, c_float_p,
                                    ctypes.c_long, c_float_p, c_float_p, ctypes.c_void_p]),
    "nerf_pack_rays": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                      c_float_p, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, ctypes.c_long,
                                      ctypes.c_int, ctypes.c_float, ctypes.c_float, c_float_p, ctypes.c_void_p]),
    "nerf_sample_coarse": (ctypes.c_int, [c_float_p, ctypes.c_long, ctypes.c_int, c_float_p, c_float_p,
                                          ctypes.c_void_p]),
    "nerf_composite_fwd": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_long,
                                          ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    "nerf_composite_bwd": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_long,
                                          ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, c_float_p,
                                          ctypes.c_void_p]),
    "nerf_composite_maps": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_long, ctypes.c_int, c_float_p,
                                           ctypes.c_void_p]),
    "nerf_to_byte": (ctypes.c_int, [c_float_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]),
    "nerf_sample_pdf": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                       c_float_p, ctypes.c_void_p]),
    "nerf_resample_merge": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_long, ctypes.c_int,
                                           ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "nerf_sample_coarse_rng": (ctypes.c_int, [c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_long,
                                              c_float_p, ctypes.c_void_p]),
    "nerf_resample_merge_rng": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_ulonglong, ctypes.c_long, ctypes.c_long,
                                               ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "nerf_composite_fwd_rng": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_float,
                                              ctypes.c_ulonglong, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                              ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    "nerf_composite_bwd_rng": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_float,
                                              ctypes.c_ulonglong, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                              ctypes.c_int, c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "nerf_rng_fill": (ctypes.c_int, [ctypes.c_int, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_long, ctypes.c_long,
                                     ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "nerf_render_scratch_bytes": (ctypes.c_size_t, [ctypes.c_long, ctypes.c_int, ctypes.c_int]),
    "nerf_render_fused": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_float_p, ctypes.c_int,
                                         ctypes.c_int, c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                         ctypes.c_ulonglong, ctypes.c_long, ctypes.c_void_p, c_float_p, c_float_p,
                                         ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "nerf_packed_model_bytes": (ctypes.c_size_t, []),
    "nerf_pack_model": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p]),
    "nerf_viewdir_term": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                         c_float_p, ctypes.c_void_p]),
    "nerf_freq_encode": (ctypes.c_int, [c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, c_float_p,
                                        ctypes.c_void_p]),
    "nerf_mlp_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int,
                                    ctypes.c_long, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p,
                                    ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]),
    "nerf_mlp_act_bytes": (ctypes.c_size_t, [ctypes.c_long]),
    "nerf_model_host_tail_bytes": (ctypes.c_size_t, []),
    "nerf_model_host_tail": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nerf_mlp_fwd_host_tail": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_float_p, c_float_p,
                                              ctypes.c_int, ctypes.c_long, ctypes.c_int, c_float_p, ctypes.c_int,
                                              c_float_p, ctypes.c_long, ctypes.c_void_p]),
    "nerf_mlp_fwd_probe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int,
                                          ctypes.c_long, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p,
                                          ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "nerf_packed_model_bwd_bytes": (ctypes.c_size_t, []),
    "nerf_pack_model_bwd": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p]),
    "nerf_pack_models_train": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                              ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                              ctypes.c_void_p]),
    "nerf_mlp_dz_bytes": (ctypes.c_size_t, [ctypes.c_long]),
    "nerf_mlp_bwd_dz": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "nerf_grad_blob_bytes": (ctypes.c_size_t, []),
    "nerf_mlp_bwd_dw": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, c_float_p, ctypes.c_void_p]),
    "nerf_mlp_bwd_heads": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_long, c_float_p, ctypes.c_void_p]),
    "nerf_viewdir_term_bwd": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                             ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "nerf_mlp_bwd_dw_det_scratch_bytes": (ctypes.c_size_t, []),
    "nerf_bwd_det_scratch_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_long]),
    "nerf_mlp_bwd_dw_det": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, c_float_p, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "nerf_mlp_bwd_heads_det": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_long, c_float_p, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "nerf_viewdir_term_bwd_det": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                                 ctypes.c_int, c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nerf_mse_loss_grad_det": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_long, c_float_p, c_float_p, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "nerf_mlp_bwd_unfold": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "nerf_grad_unpack": (ctypes.c_int, [c_float_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]),
    "nerf_mse_loss_grad": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_long, c_float_p, c_float_p,
                                          ctypes.c_void_p]),
    "nerf_adam_step": (ctypes.c_int, [ctypes.c_int] + [ctypes.POINTER(ctypes.c_void_p)] * 4 +
                       [ctypes.POINTER(ctypes.c_long), ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                        ctypes.c_long, ctypes.c_float, ctypes.c_void_p]),
    "nerf_adam_step_blob": (ctypes.c_int, [c_float_p] + [ctypes.POINTER(ctypes.c_void_p)] * 3 +
                            [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_long,
                             ctypes.c_float, ctypes.c_void_p]),
    "nerf_adam_step_blob_peers": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int] +
                                  [ctypes.POINTER(ctypes.c_void_p)] * 3 +
                                  [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_long,
                                   ctypes.c_float, ctypes.c_void_p]),
    "nerf_train_rays": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                       c_float_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_float, c_float_p, c_float_p, c_float_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
}

# symbols of the experiments build only (make -C cv-nerf_b200/csrc experiments; NERF_B200_LIB selects the
# library): the A/B variants of the field kernel behind their cycle-counter entry
_EXPERIMENT_SIGNATURES = {
    "nerf_mlp_fwd_stats": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_long, ctypes.c_int,
                                          c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
}


class NerfB200Error(RuntimeError):
    pass


def public_symbols():
    return sorted(_SIGNATURES)


def has_experiments():
    """True when the loaded library is the experiments build (libnerf_b200_exp.so)."""
    return hasattr(load(), "nerf_mlp_fwd_stats")


def load():
    """Load libnerf_b200.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfB200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in _EXPERIMENT_SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    if lib.nerf_b200_abi_version() != 1:
        raise NerfB200Error("libnerf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().nerf_b200_last_error().decode(errors="replace")
        raise NerfB200Error(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NerfB200Error("cv_nerf_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise NerfB200Error("internal: non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def f32c(t, device=None):
    """fp32, contiguous, on `device` (if given)."""
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
