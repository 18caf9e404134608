"""The train iteration of the reference (/root/reference/main.py:344-394) on the device.

Two entry levels:

* ``FusedAdam`` -- a ``torch.optim.Optimizer`` with the hyper-parameters and update rule of the
  ``torch.optim.Adam`` that ``create_model`` builds (main.py:144), stepping every tensor in one
  launch.  With it the reference's loop body (render -> loss -> backward -> optimizer.step ->
  ``param_group['lr'] = ...``) runs unchanged on the drop-in surface.

* ``TrainStep`` -- the same iteration without autograd bookkeeping: pixel batch and rays generated
  on the device (main.py:351-374), forward with saved activations, loss, compositing backward,
  field backward into two flat gradient blobs, one all-reduce of those blobs across data-parallel
  ranks (NCCL; rays are sharded, every rank draws its own batch), Adam straight from the blobs,
  re-pack of the BF16 weights, learning-rate decay (main.py:276-277, 392-394).
"""
import math
import os
import warnings

import torch

from . import kernels as K
from . import model as _model
from ._lib import NerfB200Error
from .model import Model


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas) semantics (no weight decay / amsgrad, like main.py:144)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise NerfB200Error("FusedAdam needs CUDA parameters; there is no CPU fallback")
                st = self.state[p]
                if not st:
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ps.append(p.data); gs.append(p.grad.contiguous()); ms.append(st['exp_avg']); vs.append(st['exp_avg_sq'])
            if not ps:
                continue
            group['step'] = group.get('step', 0) + 1
            K.adam_step(ps, gs, ms, vs, group['lr'], group['betas'], group['eps'], group['step'])
        _model.bump_param_epoch()
        return loss


def decayed_learning_rate(step, decay_steps, initial_lr, decay_rate=0.1):
    return initial_lr * (decay_rate ** (step / decay_steps))


def rank_seed(seed, rank):
    """Key of a rank's private random streams (pixel permutation, jitter, resampling uniforms, density
    noise): distinct for every (seed, rank) pair, so data-parallel ranks never draw the same batch."""
    return (int(seed) * 0x100000001B3 + int(rank) * 0x9E3779B1 + 1) & 0x7FFFFFFFFFFFFFFF


class TrainStep:
    """One object per process (= per GPU).  ``step(image, pose)`` runs one iteration and returns
    the loss as a 1-element device tensor (no host synchronisation)."""

    def __init__(self, coarse_model, fine_model, *, height, width, focal, n_rays=4096, n_coarse_samples=64,
                 n_fine_samples=128, perturb=1., noise=0., white_bkg=False, ndc=True, near=0., far=1.,
                 lr=5e-4, lr_decay=250, betas=(0.9, 0.999), eps=1e-8, seed=0, process_group=None):
        if not isinstance(coarse_model, Model) or not isinstance(fine_model, Model):
            raise NerfB200Error("TrainStep needs cv_nerf_b200.model.Model networks")
        self.coarse, self.fine = coarse_model, fine_model
        self.h, self.w, self.f = int(height), int(width), focal
        self.n_rays, self.s_c, self.n_fine = int(n_rays), int(n_coarse_samples), int(n_fine_samples)
        self.s_f = self.s_c + self.n_fine
        self.perturb, self.noise, self.white_bkg, self.ndc = perturb, noise, bool(white_bkg), bool(ndc)
        self.near, self.far = near, far
        self.lr0, self.lr, self.lr_decay, self.betas, self.eps = lr, lr, lr_decay, betas, eps
        self.pg = process_group
        self.world, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        # Data-parallel ranks must draw DIFFERENT ray batches (rays are sharded; identical batches
        # would make N GPUs do the work of one): the rank is folded into the pixel-permutation key and
        # into the Philox key of the stratified jitter / resampling uniforms / density noise.
        self.seed, self.it = rank_seed(seed, self.rank), 0
        dev = next(coarse_model.parameters()).device
        if dev.type != "cuda":
            raise NerfB200Error("TrainStep needs the models on a CUDA device; there is no CPU fallback")
        self.dev = dev
        g = K.grad_blob_floats()
        self.blob = torch.zeros((2, g), dtype=torch.float32, device=dev)          # [coarse, fine]
        # Data parallel: put the blobs into symmetric (peer-mapped) memory so that the fused
        # exchange+Adam kernel can read every rank's gradients over NVLink; NCCL all-reduce otherwise.
        self.symm = None
        if self.world > 1 and os.environ.get("NERF_B200_PEER_ADAM", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                group = process_group if process_group is not None else torch.distributed.group.WORLD
                blob = symm_mem.empty((2, g), dtype=torch.float32, device=dev)
                self.symm = symm_mem.rendezvous(blob, group)
                blob.zero_()
                self.blob = blob
            except Exception as exc:                      # no peer access / unsupported backend
                self.symm = None
                self.symm_error = repr(exc)
                warnings.warn(f"cv_nerf_b200.TrainStep: peer-memory gradient exchange unavailable ({exc!r}); "
                              "using the NCCL all-reduce path")
        rows_c, rows_f = self.n_rays * self.s_c, self.n_rays * self.s_f
        self.act_c = torch.empty(K.act_bytes(rows_c), dtype=torch.uint8, device=dev)
        self.act_f = torch.empty(K.act_bytes(rows_f), dtype=torch.uint8, device=dev)
        self.dz = torch.empty(K.dz_bytes(max(rows_c, rows_f)), dtype=torch.uint8, device=dev)
        self.dz_c = torch.empty(K.dz_bytes(rows_c), dtype=torch.uint8, device=dev)   # own buffer: the coarse
        # network's view-column kernel may still read it while the fine dZ chain runs
        self.params = [coarse_model.ordered_params(), fine_model.ordered_params()]
        self.m = [[torch.zeros_like(p) for p in ps] for ps in self.params]
        self.v = [[torch.zeros_like(p) for p in ps] for ps in self.params]
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.side = torch.cuda.Stream(device=dev)      # small gradient kernels next to the dW kernel
        self.profile = None                            # a list collects (start, after backward, after update) events per step
        self._early_done = False                       # forward_backward(early_update=True) has updated the coarse network

    def crop_window(self, precrop_frac=None):
        """Pre-crop window of main.py:354-361 as (row0, col0, rows, cols)."""
        if precrop_frac is None:
            return None
        dh, dw = int(self.h // 2 * precrop_frac), int(self.w // 2 * precrop_frac)
        return (self.h // 2 - dh, self.w // 2 - dw, 2 * dh, 2 * dw)

    @torch.no_grad()
    def forward_backward(self, rays, target, draws=None, early_update=False):
        """rays [n,11], target [n,3] -> loss tensor; gradients of both networks accumulated into
        self.blob (zeroed first).  ``draws``: main.RenderDraws with injected random numbers.
        ``early_update`` (used by step()): the coarse network's gradient exchange, Adam step and re-pack run
        on the side stream as soon as its gradients are complete, under the fine network's dZ / dW kernels;
        apply_gradients() then finishes with the fine network."""
        n = rays.shape[0]
        if n > self.n_rays:
            raise NerfB200Error(f"TrainStep was sized for {self.n_rays} rays per step, got {n}")
        dev = self.dev
        # Random draws: injected tensors (parity tests) or, by default, drawn inside the consuming
        # kernels from the Philox key (this rank's seed; rays of iteration `it` are numbered from
        # it * n_rays), so no torch.rand / torch.randn launch and no [n,S] tensor in HBM.
        rng = K.Rng(self.seed, self.it * self.n_rays)
        inj = lambda t: None if t is None else t.to(dev).float().contiguous()
        d_t, d_u = inj(draws.t_rand if draws else None), inj(draws.u if draws else None)
        if not self.perturb > 0.:
            z_c = K.sample_coarse(rays, self.s_c)
        elif d_t is not None:
            z_c = K.sample_coarse(rays, self.s_c, d_t)
        else:
            z_c = K.sample_coarse(rays, self.s_c, rng=rng)
        noise_c = noise_f = None
        if self.noise > 0.:
            d_nc, d_nf = inj(draws.noise_c if draws else None), inj(draws.noise_f if draws else None)
            noise_c = d_nc * self.noise if d_nc is not None else K.RngNoise(float(self.noise), rng, K.RNG_NOISE_C)
            noise_f = d_nf * self.noise if d_nf is not None else K.RngNoise(float(self.noise), rng, K.RNG_NOISE_F)

        # NERF_B200_BWD_SCHED: 2 (default) = the schedule described below; 0 / 1 = A/B timing alternatives
        sched = int(os.environ.get("NERF_B200_BWD_SCHED", "2"))
        main = torch.cuda.current_stream(dev)
        side = self.side
        pk_c, pk_f = self.coarse.packed(), self.fine.packed()
        rows_c, rows_f = n * self.s_c, n * self.s_f
        vt_c = K.viewdir_term(pk_c, rays)
        raw_c = K.mlp_fwd(pk_c, K.IN_RAYS, rays, z_c, rows_c, self.s_c, vt_c, self.s_c, act_save=self.act_c)
        rgb_c, w_c = K.composite_fwd(raw_c.view(n, self.s_c, 4), z_c, rays, noise_c, self.white_bkg)

        def loss_and_grad_raw(idx, raw, z, rgb, noise, s):
            _, g = K.mse_loss_grad(rgb, target, loss=self.loss)
            return K.composite_bwd(raw.view(n, s, 4), z, rays, noise, self.white_bkg, g)

        graw_c = None
        if sched == 2:
            # The coarse network's loss, compositing backward and head gradients need nothing from the fine
            # pass: they run on the side stream under the fine forward, as does the zeroing of the blobs.
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.loss.zero_()
                self.blob.zero_()
                graw_c = loss_and_grad_raw(0, raw_c, z_c, rgb_c, noise_c, self.s_c)
                graw_c_ready = side.record_event()
                K.mlp_bwd_heads(self.act_c, graw_c.view(rows_c, 4), rows_c, self.blob[0], stream=side)
            for t in (raw_c, rgb_c, z_c):
                t.record_stream(side)
            graw_c.record_stream(main)
        z_f = K.resample_merge(z_c, w_c, d_u) if d_u is not None else K.resample_merge(z_c, w_c, rng=rng, n_fine=self.n_fine)
        vt_f = K.viewdir_term(pk_f, rays)
        raw_f = K.mlp_fwd(pk_f, K.IN_RAYS, rays, z_f, rows_f, self.s_f, vt_f, self.s_f, act_save=self.act_f)
        if sched == 2:
            # ... and the fine network's compositing, loss and compositing backward run on the side stream under
            # the coarse dZ chain, which starts right behind the fine forward
            side.wait_stream(main)
            with torch.cuda.stream(side):
                rgb_f, _ = K.composite_fwd(raw_f.view(n, self.s_f, 4), z_f, rays, noise_f, self.white_bkg, want_weights=False)
                graw_f = loss_and_grad_raw(1, raw_f, z_f, rgb_f, noise_f, self.s_f)
                graw_f_ready = side.record_event()
                K.mlp_bwd_heads(self.act_f, graw_f.view(rows_f, 4), rows_f, self.blob[1], stream=side)
            for t in (raw_f, z_f):
                t.record_stream(side)
            graw_f.record_stream(main)
        else:
            rgb_f, _ = K.composite_fwd(raw_f.view(n, self.s_f, 4), z_f, rays, noise_f, self.white_bkg, want_weights=False)
            self.loss.zero_()
            graw_f = loss_and_grad_raw(1, raw_f, z_f, rgb_f, noise_f, self.s_f)
            graw_c = loss_and_grad_raw(0, raw_c, z_c, rgb_c, noise_c, self.s_c)
            self.blob.zero_()
            side.wait_stream(main)

        # Gradient kernels.  Main stream: dZ chain and dW per network (tensor-core kernels, one CTA per
        # SM).  Side stream: the CUDA-core kernels, placed under the kernel that leaves them room -- loss,
        # compositing backward and the l_alpha/l11 heads (inputs: saved activations + grad_raw) as above;
        # the view columns of a network follow its dZ chain.  Beside the HBM-bound dW kernel they only took
        # bandwidth from it.
        jobs = ((0, self.coarse, graw_c.view(rows_c, 4), self.act_c, rows_c, self.s_c, self.dz_c),
                (1, self.fine, graw_f.view(rows_f, 4), self.act_f, rows_f, self.s_f, self.dz))
        # NERF_B200_BWD_SCHED (A/B timing, tools/time_bwd_schedules.py): 0 everything on one stream,
        # 1 small kernels beside dW, 2 (default) as described above.  On a power-capped B200 the three
        # are within run-to-run noise of each other.  (Round 2 also tried running the two networks' dZ / dW
        # kernels side by side on disjoint SM sets -- write-bound beside read-bound: 5.05-5.58 ms against
        # 5.03-5.09 ms for this schedule, profiles/r02_bwd_schedules.txt -- and dropped it.)
        if sched == 2:
            for idx, net, graw, act, rows, s, dz in jobs:
                main.wait_event(graw_c_ready if idx == 0 else graw_f_ready)
                K.mlp_bwd_dz(net.packed_bwd(), graw, act, rows, dz=dz)
                side.wait_stream(main)
                K.viewdir_term_bwd(dz, rows, rays, s, False, self.blob[idx], stream=side)
                K.mlp_bwd_dw(act, dz, rows, self.blob[idx])
                if idx == 0:
                    # the coarse network's unfold (below) needs its dW and view-column kernels only: it runs on
                    # the side stream under the fine network's dZ chain
                    side.wait_stream(main)
                    K.mlp_bwd_unfold(self.blob[0], self.params[0], stream=side)
                    if early_update:
                        with torch.cuda.stream(side):
                            self._update_net(0, self.it + 1, True)
                        self._early_done = True
        else:
            for idx, net, graw, act, rows, s, dz in jobs:
                K.mlp_bwd_dz(net.packed_bwd(), graw, act, rows, dz=dz)
                K.mlp_bwd_params(act, dz, graw, rows, rays, s, False, self.blob[idx], side_stream=side if sched == 1 else None)
        main.wait_stream(side)
        # l9 is folded into l10 (csrc/mlp_layout.h): its gradients and those of l10's first 256 columns
        # follow from the G = dZ10^T h8 the dW kernel left in the blob and from db10 (view-column kernel)
        for idx in range(2):
            if idx == 1 or sched != 2:
                K.mlp_bwd_unfold(self.blob[idx], self.params[idx])
        return self.loss

    def _update_net(self, idx, it, allreduce):
        """Gradient exchange (data parallel), Adam from the blob and re-pack of network idx (0 coarse, 1 fine)
        on the CURRENT stream.  Peer path: barrier (all ranks' blobs of this network written), one launch that
        reads every rank's blob over NVLink, barrier (the blobs may be zeroed for the next step); every
        network has its own pair of barrier channels, so the two updates may run on different streams."""
        dev = self.dev
        if self.world > 1 and allreduce and self.symm is not None:
            stream = torch.cuda.current_stream(dev).cuda_stream
            g_bytes = self.blob.shape[1] * 4
            self.symm.barrier(channel=2 * idx)
            peers = [int(ptr) + idx * g_bytes for ptr in self.symm.buffer_ptrs]
            K.adam_step_blob_peers(peers, [p.data for p in self.params[idx]], self.m[idx], self.v[idx], self.lr,
                                   self.betas, self.eps, it, 1. / self.world, stream)
            self.symm.barrier(channel=2 * idx + 1)
        else:
            if self.world > 1 and allreduce:
                torch.distributed.all_reduce(self.blob[idx], group=self.pg)
            # without a reduction the blob holds this rank's own mean gradient: no 1/world
            K.adam_step_blob(self.blob[idx], [p.data for p in self.params[idx]], self.m[idx], self.v[idx],
                             self.lr, self.betas, self.eps, it,
                             grad_scale=1. / self.world if allreduce else 1.)
        net = self.coarse if idx == 0 else self.fine
        buf = net.packed_buffers()
        K.pack_models_train([self.params[idx]], [buf[0]], [buf[1]])

    @torch.no_grad()
    def apply_gradients(self, allreduce=True):
        """all-reduce (data parallel), Adam from the blobs, weight re-pack, learning-rate decay -- for the
        networks forward_backward(early_update=True) has not already updated."""
        self.it += 1
        early, self._early_done = self._early_done, False
        if early or (self.world > 1 and allreduce):
            for idx in ((1,) if early else (0, 1)):
                self._update_net(idx, self.it, allreduce)
        else:
            for idx in range(2):
                K.adam_step_blob(self.blob[idx], [p.data for p in self.params[idx]], self.m[idx], self.v[idx],
                                 self.lr, self.betas, self.eps, self.it, grad_scale=1.)
            # re-pack both networks' forward and transposed blobs in one launch
            bufs = [net.packed_buffers() for net in (self.coarse, self.fine)]
            K.pack_models_train(self.params, [b[0] for b in bufs], [b[1] for b in bufs])
        _model.bump_param_epoch()
        self.coarse.mark_packed(); self.fine.mark_packed()
        # main.py:392-394: the decayed rate takes effect from the next iteration on
        self.lr = decayed_learning_rate(self.it, self.lr_decay * 1000, self.lr0)

    @torch.no_grad()
    def step(self, image, pose, *, precrop_frac=None, pix=None, draws=None):
        rays, target, _ = K.train_rays(self.h, self.w, self.f, pose, self.n_rays, pix=pix,
                                       seed=self.seed * 0x9E3779B97F4A7C15 + self.it, crop=self.crop_window(precrop_frac),
                                       image=image, ndc=self.ndc, near=self.near, far=self.far)
        prof = self.profile
        if prof is not None:        # bench.py: CUDA events around the two halves of the step (no synchronisation here)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        loss = self.forward_backward(rays, target, draws, early_update=True)
        if prof is not None:
            ev[1].record()
        self.apply_gradients()
        if prof is not None:
            ev[2].record()
            prof.append(ev)
        return loss

    def gradients(self, idx):
        """The current blob of network idx (0 coarse, 1 fine) unpacked into 24 tensors."""
        grads = [torch.empty_like(p) for p in self.params[idx]]
        return K.grad_unpack(self.blob[idx], grads)


# ------------------------------------------------------------------ checkpoints
def save_checkpoint(path, step, coarse_model, fine_model, optimizer=None, train_step=None):
    """{step}.pt in the layout NeRF trainers of this family write (the reference's own
    ``results/lego/{N}.pt`` files are listed in its .MISSING_LARGE_BLOBS but the saving code is
    gone): state_dicts with the reference's parameter names (model.py:57-71)."""
    blob = {"global_step": int(step), "coarse_state_dict": coarse_model.state_dict(),
            "fine_state_dict": fine_model.state_dict()}
    if optimizer is not None:
        blob["optimizer_state_dict"] = optimizer.state_dict()
    if train_step is not None:
        blob["train_step"] = {"it": train_step.it, "lr": train_step.lr,
                              "m": [[t.cpu() for t in ms] for ms in train_step.m],
                              "v": [[t.cpu() for t in vs] for vs in train_step.v]}
    torch.save(blob, path)


def load_checkpoint(path, coarse_model, fine_model, optimizer=None, train_step=None, map_location=None):
    blob = torch.load(path, map_location=map_location, weights_only=True)   # tensors, numbers, lists, dicts only
    coarse_model.load_state_dict(blob["coarse_state_dict"])
    fine_model.load_state_dict(blob["fine_state_dict"])
    _model.bump_param_epoch()
    if optimizer is not None and "optimizer_state_dict" in blob:
        optimizer.load_state_dict(blob["optimizer_state_dict"])
    if train_step is not None and "train_step" in blob:
        st = blob["train_step"]
        train_step.it, train_step.lr = st["it"], st["lr"]
        for dst, src in zip(train_step.m + train_step.v, st["m"] + st["v"]):
            for d, s_ in zip(dst, src):
                d.copy_(s_)
    return blob["global_step"]
