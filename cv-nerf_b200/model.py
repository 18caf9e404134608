"""Drop-in for the reference's ``model.py``: FreqEmbedding, Model, net_forward, combine, to_byte.

``Model`` keeps the reference's parameter names, shapes and construction order
(/root/reference/model.py:57-71), so ``state_dict`` round-trips and ``torch.manual_seed(s)``
yields the same initial weights.  Its arithmetic runs in the fused sm_100a kernel
(csrc/mlp_fwd.cu): BF16 tensor-core contractions with FP32 accumulation, FP32 sigma/rgb heads.
"""
import numpy as np
import torch
import torch.nn as nn

from . import kernels as K
from ._lib import NerfB200Error, f32c

def _lib_bwd_bytes():
    from . import _lib
    return _lib.load().nerf_packed_model_bwd_bytes()


STD_CHUNK_SIZE = 65536
_PARAM_EPOCH = 0


def bump_param_epoch():
    """Called by in-place parameter updates that bypass torch's version counters (the fused Adam
    kernels): invalidates every Model's packed-weight cache."""
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1

PARAM_ORDER = ("l1", "l2", "l3", "l4", "l5", "l6", "l7", "l8", "l9", "l_alpha", "l10", "l11")


class FreqEmbedding:
    """Positional encoding (model.py:9-31): [x, sin(2^k x), cos(2^k x)]_{k<freqs}.

    The fused render path never materialises this tensor (the field kernel encodes points
    itself); ``embed`` serves callers of the reference's ``Model.forward(x[...,90])`` interface
    and runs the stand-alone encoder kernel (csrc/mlp_pack.cu, nerf_freq_encode)."""

    def __init__(self, freqs, dim=3):
        self.freqs = freqs
        self.dim = dim
        self.out_dim = dim + 2 * dim * freqs

    def embed(self, inputs):
        if not inputs.is_cuda:
            raise NerfB200Error("FreqEmbedding.embed needs a CUDA tensor; there is no CPU fallback")
        return K.freq_encode(inputs, self.freqs)


class _FieldFn(torch.autograd.Function):
    """raw = field(inputs) with parameter gradients.  ``spec`` is a dict describing the input
    mode (see kernels.mlp_fwd); the parameters follow as individual tensors so autograd tracks
    them."""

    @staticmethod
    def forward(ctx, model, spec, *params):
        act = torch.empty(K.act_bytes(spec["rows"]), dtype=torch.uint8, device=spec["in0"].device)
        raw = model._forward_raw(spec, act_save=act)
        ctx.model, ctx.spec, ctx.act = model, spec, act
        # the transposed weights must be the ones this forward used (an optimizer step may run
        # before backward in exotic loops); packing is one small kernel
        ctx.packed_bwd = model.packed_bwd()
        return raw

    @staticmethod
    def backward(ctx, grad_raw):
        grads = ctx.model._backward_raw(ctx.spec, ctx.act, ctx.packed_bwd, grad_raw.contiguous())
        ctx.act = None
        return (None, None) + tuple(grads)


class Model(nn.Module):
    def __init__(self, xyz_L=10, angle_L=4):
        super().__init__()
        if (xyz_L, angle_L) != (10, 4):
            raise NerfB200Error("the fused field kernel is built for xyz_L=10, angle_L=4 "
                                "(the only configuration the reference uses, main.py:129-136)")
        self.xyz_L, self.angle_L = xyz_L, angle_L
        enc = self._encoding_dim(3, xyz_L) + 3
        enc_dir = self._encoding_dim(3, angle_L) + 3
        # construction order == reference (RNG parity of the default init)
        self.l1 = nn.Linear(enc, 256)
        self.l2 = nn.Linear(256, 256)
        self.l3 = nn.Linear(256, 256)
        self.l4 = nn.Linear(256, 256)
        self.l5 = nn.Linear(256, 256)
        self.l6 = nn.Linear(256 + enc, 256)
        self.l7 = nn.Linear(256, 256)
        self.l8 = nn.Linear(256, 256)
        self.l9 = nn.Linear(256, 256)
        self.l_alpha = nn.Linear(256, 1)
        self.l10 = nn.Linear(256 + enc_dir, 128)
        self.l11 = nn.Linear(128, 3)
        self._packed = None
        self._packed_key = None
        self._packed_bwd = None
        self._packed_bwd_key = None
        self._host_tail = None
        self._host_tail_key = None

    @staticmethod
    def _encoding_dim(num_comp, L):
        return 2 * num_comp * L

    # ------------------------------------------------------------------ packed parameters
    def ordered_params(self):
        out = []
        for name in PARAM_ORDER:
            lin = getattr(self, name)
            out += [lin.weight, lin.bias]
        return out

    def packed(self):
        """BF16 UMMA-layout blob of the current parameters; re-packed (one small kernel) when any
        parameter was modified in place (optimizer step, load_state_dict) or moved."""
        params = self.ordered_params()
        key = (_PARAM_EPOCH,) + tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or key != self._packed_key:
            if not params[0].is_cuda:
                raise NerfB200Error("Model parameters must live on a CUDA device; there is no CPU fallback")
            self._packed = K.pack_model(params, self._packed if self._packed is not None
                                        and self._packed.device == params[0].device else None)
            self._packed_key = key
        return self._packed

    def _cache_key(self):
        return (_PARAM_EPOCH,) + tuple((p.data_ptr(), p._version) for p in self.ordered_params())

    def packed_buffers(self):
        """(packed, packed_bwd) buffers, allocated if needed, WITHOUT re-packing (for
        kernels.pack_models_train, which fills several Models' blobs in one launch)."""
        dev = self.l1.weight.device
        if self._packed is None or self._packed.device != dev:
            self._packed = torch.empty(K.packed_model_bytes(), dtype=torch.uint8, device=dev)
        if self._packed_bwd is None or self._packed_bwd.device != dev:
            self._packed_bwd = torch.empty(int(_lib_bwd_bytes()), dtype=torch.uint8, device=dev)
        return self._packed, self._packed_bwd

    def mark_packed(self):
        """The blobs returned by packed_buffers() now hold the current parameters."""
        self._packed_key = self._packed_bwd_key = self._cache_key()

    def host_tail(self):
        """Host copy of the biases and heads for the inference kernel variant (one stream
        synchronisation per weight change; the training path never calls this)."""
        packed = self.packed()
        if self._host_tail is None or self._host_tail_key != self._packed_key:
            self._host_tail = K.model_host_tail(packed)
            self._host_tail_key = self._packed_key
        return self._host_tail

    def packed_bwd(self):
        """Transposed BF16 weights for the backward dZ chain (same caching rule as packed())."""
        params = self.ordered_params()
        key = (_PARAM_EPOCH,) + tuple((p.data_ptr(), p._version) for p in params)
        if self._packed_bwd is None or key != self._packed_bwd_key:
            if not params[0].is_cuda:
                raise NerfB200Error("Model parameters must live on a CUDA device; there is no CPU fallback")
            self._packed_bwd = K.pack_model_bwd(params)   # fresh buffer: an in-flight backward may hold the old one
            self._packed_bwd_key = key
        return self._packed_bwd

    # ------------------------------------------------------------------ kernels
    def _forward_raw(self, spec, act_save=None):
        packed = self.packed()
        vterm = K.viewdir_term(packed, spec["dirs"], embedded=spec.get("dirs_embedded", False))
        return K.mlp_fwd(packed, spec["mode"], spec["in0"], spec.get("in1"), spec["rows"],
                         spec.get("samples", 1), vterm, spec["vterm_div"], spec.get("in_stride", 0),
                         act_save=act_save, host_tail=None if act_save is not None else self.host_tail(),
                         row0=spec.get("row0", 0))

    def _backward_blob(self, spec, act, packed_bwd, grad_raw, blob):
        """grad_raw [rows,4] -> parameter gradients accumulated into the fp32 blob (see
        csrc/mlp_bwd_layout.h): dZ chain, then dW/db contractions."""
        rows = spec["rows"]
        dz = K.mlp_bwd_dz(packed_bwd, grad_raw.reshape(rows, 4), act, rows)
        K.mlp_bwd_params(act, dz, grad_raw.reshape(rows, 4), rows, spec["dirs"], spec["vterm_div"],
                         spec.get("dirs_embedded", False), blob, params=self.ordered_params())
        return blob

    def _backward_raw(self, spec, act, packed_bwd, grad_raw):
        dev = grad_raw.device
        blob = torch.zeros(K.grad_blob_floats(), dtype=torch.float32, device=dev)
        self._backward_blob(spec, act, packed_bwd, grad_raw, blob)
        grads = [torch.empty_like(p) for p in self.ordered_params()]
        K.grad_unpack(blob, grads)
        # autograd wants them in the order the parameters were passed: ordered_params()
        return grads

    def field(self, spec):
        """Evaluate the network for an input spec; differentiable w.r.t. the parameters."""
        params = self.ordered_params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _FieldFn.apply(self, spec, *params)
        return self._forward_raw(spec)

    def forward(self, x):
        """x [..., 90] = [PE10(point) | PE4(viewdir)] -> [..., 4] = (rgb_raw, sigma_raw)
        (model.py:77-107)."""
        if x.shape[-1] != 90:
            raise NerfB200Error(f"Model.forward expects 90 input columns, got {x.shape[-1]}")
        flat = f32c(x.reshape(-1, 90))
        if not flat.is_cuda:
            raise NerfB200Error("Model.forward needs a CUDA tensor; there is no CPU fallback")
        rows = flat.shape[0]
        spec = dict(mode=K.IN_EMBEDDED, in0=flat, rows=rows, in_stride=90, vterm_div=1,
                    dirs=flat[:, 63:], dirs_embedded=True)
        return self.field(spec).reshape(*x.shape[:-1], 4)


def net_forward(inputs, dirs, f, embed_fn=None, embeddirs_fn=None, netchunk=1024 * 64):
    """inputs [n,S,3], dirs [n,3] -> [n,S,4] (model.py:110-122).  ``embed_fn`` / ``embeddirs_fn``
    / ``netchunk`` are accepted for signature compatibility: the fused kernel encodes points and
    view directions itself and never materialises the [n*S,90] tensor."""
    if not isinstance(f, Model):
        raise NerfB200Error("net_forward needs a cv_nerf_b200.model.Model; there is no fallback path")
    if dirs is None:
        raise NerfB200Error("net_forward without view directions is not supported (the reference "
                            "always passes them, main.py:219)")
    pts = f32c(inputs.reshape(-1, 3))
    n, s = inputs.shape[0], inputs.shape[1]
    spec = dict(mode=K.IN_POINTS, in0=pts, rows=pts.shape[0], vterm_div=s, dirs=f32c(dirs))
    return f.field(spec).reshape(list(inputs.shape[:-1]) + [4])


def combine(f, chunk):
    """model.py:125-131: chunked application of f along dim 0."""
    if chunk is None:
        return f

    def ret(inputs):
        return torch.cat([f(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def to_byte(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)
