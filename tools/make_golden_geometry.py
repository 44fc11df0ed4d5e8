"""Goldens for the geometry helpers from the REFERENCE functions run on CPU here:
BEVFormerEncoder.get_reference_points / point_sampling (modules/encoder.py:53-156).
Writes tests/golden/geometry.npz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import geometry_cases as gc  # noqa: E402
from tools import ref_shim  # noqa: E402


def main():
    enc = ref_shim.load("modules.encoder").BEVFormerEncoder
    H, W = gc.BEV
    ref3d = enc.get_reference_points(H, W, 8, 4, dim="3d", bs=2, device="cpu", dtype=torch.float)
    ref2d = enc.get_reference_points(H, W, dim="2d", bs=2, device="cpu", dtype=torch.float)
    cam, mask = enc.point_sampling(None, ref3d, gc.PC_RANGE, gc.rig(2))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "geometry.npz"), ref3d=ref3d.numpy(), ref2d=ref2d.numpy(),
                        ref_cam=cam.numpy(), bev_mask=mask.numpy())
    print(ref3d.shape, ref2d.shape, cam.shape, mask.shape, float(mask.float().mean()))


if __name__ == "__main__":
    main()
