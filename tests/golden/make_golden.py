"""Generate golden fixtures by running the REAL reference (read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
It imports the reference with the stub recipe of SURVEY.md section 8(c), injects
recorded random draws by patching torch.rand / torch.randn for the duration of
each call, and writes small .npz fixtures next to this file.  Nothing from the
reference's sources is copied; only its numerical outputs are stored.
"""
import hashlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    for name in ["imageio", "configargparse", "skimage", "skimage.transform", "skimage.io", "skimage.color"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.transform"].rescale = None
    sys.modules["skimage.io"].imread = None
    sys.modules["skimage.color"].rgba2rgb = None
    sys.path.insert(0, REF)
    import main as ref_main            # noqa
    import data_helpers as ref_dh      # noqa
    ref_main.batchify_rays = ref_main.batch_rays      # main.py:79 calls an undefined name
    torch.autograd.set_detect_anomaly(False)
    return ref_main, ref_dh


class InjectDraws:
    """Make torch.rand / torch.randn hand out pre-recorded tensors, in order."""

    def __init__(self, rand_list, randn_list):
        self.rand_list, self.randn_list = list(rand_list), list(randn_list)

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def fake_rand(*shape, **kw):
            t = self.rand_list.pop(0)
            want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(t.shape) == want, (t.shape, want)
            return t

        def fake_randn(*shape, **kw):
            t = self.randn_list.pop(0)
            want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(t.shape) == want, (t.shape, want)
            return t

        torch.rand, torch.randn = fake_rand, fake_randn
        return self

    def __exit__(self, *a):
        torch.rand, torch.randn = self._rand, self._randn
        assert not self.rand_list and not self.randn_list, "unused draws"


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def weights_digest(models) -> str:
    h = hashlib.sha256()
    for m in models:
        for k, v in m.state_dict().items():
            h.update(k.encode())
            h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_models(ref_main, seed, sigma_bias, sigma_gain, **flags):
    torch.manual_seed(seed)
    args = SimpleNamespace(netchunk=65536, lr=5e-4, perturb=flags.get("perturb", 0.),
                           n_fine_samples=flags.get("n_fine", 128), n_coarse_samples=64,
                           white_bkg=flags.get("white_bkg", False), noise=flags.get("noise", 0.),
                           dtype=flags.get("dtype", "blender"), no_ndc=flags.get("no_ndc", False))
    kw_train, kw_test, _, grad_vars, _ = ref_main.create_model(args)
    for m in (kw_train["coarse_model"], kw_train["fine_model"]):
        if sigma_bias is not None:
            with torch.no_grad():
                m.l_alpha.bias.fill_(sigma_bias)
                m.l_alpha.weight.mul_(sigma_gain)
    return kw_train, kw_test, grad_vars


def golden_rays(ref_main, ref_dh):
    out = {}
    cases = {
        "lego400": (400, 400, 555.5555155968841, ref_dh.pose_spherical(-180., -30., 4.)[:3, :4]),
        "lego800": (800, 800, 1111.1110311937682, ref_dh.pose_spherical(36., -30., 4.)[:3, :4]),
        "fern": (378, 504, np.float32(407.5657), None),
    }
    g = torch.Generator().manual_seed(7)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    fern_pose = torch.cat([q, torch.tensor([[0.31], [-0.12], [0.05]])], 1).float()
    for name, (h, w, f, pose) in cases.items():
        pose = fern_pose if pose is None else pose.float()
        o, d = ref_main.compute_rays(h, w, f, pose)
        out[f"{name}_hwf"] = np.array([h, w, float(f)], dtype=np.float64)
        out[f"{name}_f_is_f32"] = np.array(isinstance(f, np.float32))
        out[f"{name}_pose"] = pose.numpy()
        out[f"{name}_d_sha"] = np.array(sha(d))
        out[f"{name}_o_sha"] = np.array(sha(o.contiguous()))
        out[f"{name}_d_sub"] = d[::37, ::41].contiguous().numpy()
        if name == "fern":
            on, dn = ref_dh.get_ndc(h, w, f, 1., o, d)
            out["fern_ndc_o_sha"] = np.array(sha(on))
            out["fern_ndc_d_sha"] = np.array(sha(dn))
            out["fern_ndc_o_sub"] = on[::37, ::41].contiguous().numpy()
            out["fern_ndc_d_sub"] = dn[::37, ::41].contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "rays.npz"), **out)
    print("rays.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def golden_units(ref_main):
    """Stand-alone compositing / sampler / encoder vectors."""
    import model as ref_model
    import utils as ref_utils
    g = torch.Generator().manual_seed(11)
    n = 48
    out = {}
    for s_len, tag in ((64, "c"), (192, "f")):
        raw = torch.randn(n, s_len, 4, generator=g) * 2.
        z, _ = torch.sort(2. + 4. * torch.rand(n, s_len, generator=g), -1)
        d = torch.randn(n, 3, generator=g)
        noise_draw = torch.randn(n, s_len, generator=g)
        rgb0, w0 = ref_main.process_volume_info(raw, z, d, 0., True)
        with InjectDraws([], [noise_draw]):
            rgb1, w1 = ref_main.process_volume_info(raw, z, d, 0.7, False)
        out.update({f"comp_{tag}_raw": raw.numpy(), f"comp_{tag}_z": z.numpy(), f"comp_{tag}_d": d.numpy(),
                    f"comp_{tag}_noise_draw": noise_draw.numpy(),
                    f"comp_{tag}_rgb_white": rgb0.numpy(), f"comp_{tag}_w_white": w0.numpy(),
                    f"comp_{tag}_rgb_noise": rgb1.numpy(), f"comp_{tag}_w_noise": w1.numpy()})
    bins, _ = torch.sort(2. + 4. * torch.rand(n, 63, generator=g), -1)
    wts = torch.rand(n, 62, generator=g) ** 4
    wts[3] = 0.                    # all-epsilon pdf
    wts[5, :] = 0.; wts[5, 17] = 1.  # a spike
    u = torch.rand(n, 128, generator=g)
    u[0, 0] = 0.; u[0, 1] = 1. - 2 ** -24
    with InjectDraws([u], []):
        smp = ref_utils.inv_transform_sampling(bins, wts, 128)
    out.update(pdf_bins=bins.numpy(), pdf_w=wts.numpy(), pdf_u=u.numpy(), pdf_samples=smp.numpy())
    x = torch.randn(n, 3, generator=g) * 3.
    x[0] = torch.tensor([1e8, -2.3e8, 0.])
    out.update(enc_x=x.numpy(), enc10=ref_model.FreqEmbedding(10).embed(x).numpy(),
               enc4=ref_model.FreqEmbedding(4).embed(x).numpy())
    np.savez_compressed(os.path.join(HERE, "units.npz"), **out)
    print("units.npz written")


def golden_render(ref_main, ref_dh, name, *, h, w, f, pose, pix, seed, sigma_bias, sigma_gain,
                  white_bkg, ndc, near, far, train):
    flags = dict(white_bkg=white_bkg, dtype="llff" if ndc else "blender",
                 perturb=1. if train else 0., noise=1. if (train and ndc) else 0.)
    kw_train, kw_test, grad_vars = make_models(ref_main, seed, sigma_bias, sigma_gain, **flags)
    kw = dict(kw_train if train else kw_test)
    kw.update(near=near, far=far)
    n = pix.shape[0]
    g = torch.Generator().manual_seed(seed + 100)
    o_full, d_full = ref_main.compute_rays(h, w, f, pose)
    rays = torch.stack([o_full[pix[:, 0], pix[:, 1]], d_full[pix[:, 0], pix[:, 1]]], 0)
    u = torch.rand(n, 128, generator=g)
    rand_list, randn_list = [], []
    out = {}
    if train:
        t_rand = torch.rand(n, 64, generator=g)
        rand_list.append(t_rand)
        out["t_rand"] = t_rand.numpy()
        if kw["noise"] > 0:
            nc, nf = torch.randn(n, 64, generator=g), torch.randn(n, 192, generator=g)
            randn_list += [nc, nf]
            out["noise_c"], out["noise_f"] = nc.numpy(), nf.numpy()
    rand_list.append(u)
    target = torch.rand(n, 3, generator=g)
    with InjectDraws(rand_list, randn_list):
        rgb, extras = ref_main.render(h, w, f, chunk=32768, rays=rays, **kw)
    out.update(hwf=np.array([h, w, float(f)]), f_is_f32=np.array(isinstance(f, np.float32)),
               pose=pose.numpy(), pix=pix.numpy(), seed=np.array(seed),
               sigma_bias=np.array(np.nan if sigma_bias is None else sigma_bias), sigma_gain=np.array(sigma_gain),
               white_bkg=np.array(white_bkg), ndc=np.array(ndc), near=np.array(near), far=np.array(far),
               train=np.array(train), noise=np.array(kw["noise"]), u=u.numpy(), target=target.numpy(),
               rays_o=rays[0].numpy(), rays_d=rays[1].numpy(),
               rgb_map=rgb.detach().numpy(), rgb_c=extras["rgb_c"].detach().numpy(),
               weights_sha=np.array(weights_digest([kw["coarse_model"], kw["fine_model"]])))
    if train:
        loss = torch.mean((rgb - target) ** 2) + torch.mean((extras["rgb_c"] - target) ** 2)
        loss.backward()
        out["loss"] = np.array(loss.item())
        names = [f"{m}.{k}" for m in ("coarse", "fine")
                 for k, _ in kw["coarse_model"].named_parameters()]
        for nm, p in zip(names, grad_vars):
            gr = p.grad.detach().reshape(-1)
            out[f"gnorm/{nm}"] = np.array(gr.norm().item())
            out[f"ghead/{nm}"] = gr[:96].numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), **out)
    print(f"render_{name}.npz: rgb mean {rgb.mean().item():.4f} rgb_c mean {extras['rgb_c'].mean().item():.4f}")


def main():
    ref_main, ref_dh = import_reference()
    golden_rays(ref_main, ref_dh)
    golden_units(ref_main)
    g = torch.Generator().manual_seed(3)
    pix = torch.stack([torch.randint(100, 300, (96,), generator=g), torch.randint(100, 300, (96,), generator=g)], 1)
    lego = dict(h=400, w=400, f=555.5555155968841, pose=ref_dh.pose_spherical(-180., -30., 4.)[:3, :4].float(),
                pix=pix, white_bkg=True, ndc=False, near=2., far=6.)
    golden_render(ref_main, ref_dh, "lego_test", seed=0, sigma_bias=1.0, sigma_gain=5.0, train=False, **lego)
    golden_render(ref_main, ref_dh, "lego_test_stock", seed=0, sigma_bias=None, sigma_gain=1.0, train=False, **lego)
    golden_render(ref_main, ref_dh, "lego_train", seed=1, sigma_bias=0.5, sigma_gain=20.0, train=True, **lego)
    pixf = torch.stack([torch.randint(0, 378, (64,), generator=g), torch.randint(0, 504, (64,), generator=g)], 1)
    fern_pose = torch.tensor([[0.998, 0.02, -0.05, 0.12], [-0.018, 0.999, 0.03, -0.07],
                              [0.051, -0.029, 0.998, 0.35]], dtype=torch.float32)
    fern = dict(h=378, w=504, f=np.float32(407.5657), pose=fern_pose, pix=pixf,
                white_bkg=False, ndc=True, near=0., far=1.)
    golden_render(ref_main, ref_dh, "fern_test", seed=2, sigma_bias=1.0, sigma_gain=5.0, train=False, **fern)
    golden_render(ref_main, ref_dh, "fern_train", seed=2, sigma_bias=1.0, sigma_gain=5.0, train=True, **fern)


if __name__ == "__main__":
    main()
