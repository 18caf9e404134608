"""torch.autograd glue: forward/backward of the two differentiable stages of the render path
(field network, compositing) as Functions over the C-ABI kernels, so ``loss.backward()`` and
``optimizer.step()`` of the reference's train loop (main.py:379-386) work unchanged."""
import torch

from . import kernels as K


class CompositeFn(torch.autograd.Function):
    """process_volume_info (main.py:174-204); differentiable w.r.t. raw only (z, dirs and the
    noise draw have no learnable ancestor in the reference either)."""

    @staticmethod
    def forward(ctx, raw, z, dirs, noise, white_bkg):
        n, s = z.shape
        rgb, w = K.composite_fwd(raw.reshape(n, s, 4), z, dirs, noise, white_bkg)
        # noise: None, a tensor, or a kernels.RngNoise key (the backward regenerates the same draws)
        is_tensor = isinstance(noise, torch.Tensor)
        ctx.save_for_backward(raw, z, dirs, noise if is_tensor else None)
        ctx.noise_key = None if is_tensor else noise
        ctx.white_bkg = white_bkg
        ctx.set_materialize_grads(False)
        return rgb, w

    @staticmethod
    def backward(ctx, grad_rgb, grad_w):
        raw, z, dirs, noise = ctx.saved_tensors
        noise = ctx.noise_key if noise is None else noise
        n, s = z.shape
        if grad_rgb is None and grad_w is None:
            return None, None, None, None, None
        if grad_rgb is None:
            grad_rgb = torch.zeros((n, 3), dtype=torch.float32, device=raw.device)
        gw = None if grad_w is None else grad_w.contiguous()
        g = K.composite_bwd(raw.reshape(n, s, 4), z, dirs, noise, ctx.white_bkg, grad_rgb.contiguous(), gw)
        return g.reshape(raw.shape), None, None, None, None


def composite(raw, z, dirs, noise, white_bkg, want_weights=True):
    if torch.is_grad_enabled() and raw.requires_grad:
        return CompositeFn.apply(raw, z, dirs, noise, white_bkg)
    n, s = z.shape
    return K.composite_fwd(raw.reshape(n, s, 4), z, dirs, noise, white_bkg, want_weights=want_weights)
