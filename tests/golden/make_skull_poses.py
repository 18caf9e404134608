"""Derive the skull capture's camera poses with the REFERENCE's own pose math
(/root/reference/data_helpers.py:261-324 load_llff_data: axis fix-up, rescale by bounds, recenter,
120-pose spiral) from /root/reference/skull/poses_bounds.npy, with only the image read stubbed
(the skull images are not in the tree).  Build-container only; writes skull_spiral.npz.

BASELINE.json configs[4] ("configs/skull.txt LLFF spiral path, 120-frame novel-view video") renders
these poses; nothing of the reference's source is copied, only the poses it computes."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

FACTOR = 8   # configs/skull.txt


def main():
    ref_main, ref_dh = import_reference()
    topdir = "/root/reference/skull"

    def load_llff_without_images(_topdir, factor=None):
        pb = np.load(os.path.join(topdir, "poses_bounds.npy"))
        poses = pb[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0])          # data_helpers.py:126-127
        bounds = pb[:, -2:].transpose([1, 0])                                # :130-131
        h, w = int(poses[0, 4, 0]) // factor, int(poses[1, 4, 0]) // factor   # what cv2.resize would produce
        images = np.zeros((h, w, 3, poses.shape[-1]), dtype=np.float32)
        poses[:2, 4, :] = np.array([h, w]).reshape([2, 1])                    # :188-190
        poses[2, 4, :] = poses[2, 4, :] * 1. / factor
        return poses, bounds, images

    ref_dh.load_llff = load_llff_without_images
    images, pose, render_poses, hwf, i_test, bounds = ref_dh.load_llff_data(topdir, factor=FACTOR)
    render_poses = np.stack(render_poses, 0).astype(np.float32)
    out = dict(hwf=np.asarray(hwf, dtype=np.float32), train_poses=pose.astype(np.float32),
               render_poses=render_poses[:, :3, :4], bounds=bounds.astype(np.float32), i_test=np.array(i_test))
    np.savez_compressed(os.path.join(HERE, "skull_spiral.npz"), **out)
    print("hwf", out["hwf"], "render poses", out["render_poses"].shape, "train poses", out["train_poses"].shape,
          "bounds", float(bounds.min()), float(bounds.max()))


if __name__ == "__main__":
    main()
