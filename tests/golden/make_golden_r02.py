"""Round-2 fixtures from the REAL reference (build container only; same recipe as make_golden.py,
whose helpers are reused): the configurations BASELINE.json quotes the metric on.

    python tests/golden/make_golden_r02.py

* render_lego800_test.npz     lego full-res 800x800 (configs[1]), 128 pixels of the frame
* render_skull_f{000,030,060,090}_test.npz
                              configs/skull.txt: the reference's own spiral poses (skull_spiral.npz,
                              derived by make_skull_poses.py) through render(..., ndc=True); frames 0 and
                              60 have NDC origins of ~2.3e8 -- they are what the
                              reference renders, so they are kept.
Nothing of the reference's source is copied; only numerical outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import golden_render, import_reference  # noqa: E402


def main():
    ref_main, ref_dh = import_reference()
    g = torch.Generator().manual_seed(21)
    pix = torch.stack([torch.randint(0, 800, (128,), generator=g), torch.randint(0, 800, (128,), generator=g)], 1)
    lego = dict(h=800, w=800, f=1111.1110311937682, pose=ref_dh.pose_spherical(36., -30., 4.)[:3, :4].float(),
                pix=pix, white_bkg=True, ndc=False, near=2., far=6.)
    golden_render(ref_main, ref_dh, "lego800_test", seed=0, sigma_bias=1.0, sigma_gain=5.0, train=False, **lego)

    sk = np.load(os.path.join(HERE, "skull_spiral.npz"))
    h, w, f = int(sk["hwf"][0]), int(sk["hwf"][1]), np.float32(sk["hwf"][2])
    for frame in (0, 30, 60, 90):
        pixs = torch.stack([torch.randint(0, h, (96,), generator=g), torch.randint(0, w, (96,), generator=g)], 1)
        pose = torch.from_numpy(sk["render_poses"][frame]).float()
        golden_render(ref_main, ref_dh, f"skull_f{frame:03d}_test", seed=2, sigma_bias=1.0, sigma_gain=5.0,
                      train=False, h=h, w=w, f=f, pose=pose, pix=pixs, white_bkg=False, ndc=True, near=0., far=1.)


if __name__ == "__main__":
    main()
