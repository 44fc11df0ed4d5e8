"""Geometry helpers against goldens from the reference functions (tools/make_golden_geometry.py).
Grid generation runs on CPU; the CUDA point_sampling kernel is a gpu test."""
import os

import numpy as np
import pytest
import torch

from tests import geometry_cases as gc
from vidar_b200 import bev_geometry as geo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "geometry.npz")


def test_reference_point_grids_match_reference():
    g = np.load(GOLD)
    H, W = gc.BEV
    np.testing.assert_array_equal(geo.get_reference_points(H, W, 8, 4, "3d", 2, "cpu").numpy(), g["ref3d"])
    np.testing.assert_array_equal(geo.get_reference_points(H, W, dim="2d", bs=2, device="cpu").numpy(), g["ref2d"])
    np.testing.assert_array_equal(geo.get_bev_grids(H, W, 2, "cpu").numpy(), g["ref2d"][:, :, 0])
    pts = torch.tensor([[[0.0, 0.0, -1.0], [51.2, -51.2, 3.0]]])
    vox = geo.coords_to_voxel_grids(pts, 200, 200, 16, gc.PC_RANGE)
    np.testing.assert_allclose(vox.numpy(), [[[100, 100, 8], [200, 0, 16]]], rtol=1e-6)
    assert torch.equal(pts, torch.tensor([[[0.0, 0.0, -1.0], [51.2, -51.2, 3.0]]]))     # input untouched


@pytest.mark.gpu
def test_point_sampling_kernel_matches_reference(cuda):
    g = np.load(GOLD)
    ref3d = torch.from_numpy(g["ref3d"]).to(cuda)
    cam, mask = geo.point_sampling(ref3d, gc.PC_RANGE, gc.rig(2))
    assert cam.shape == g["ref_cam"].shape and mask.shape == g["bev_mask"].shape
    # masked-in points: coordinates to 1e-5; the mask itself may differ only for points within
    # rounding distance of the image border / the z = eps plane
    same = mask.cpu().numpy() == g["bev_mask"]
    assert same.mean() > 0.999
    keep = g["bev_mask"] & same
    np.testing.assert_allclose(cam.cpu().numpy()[keep], g["ref_cam"][keep], rtol=1e-4, atol=1e-5)
    # everything in front of the camera agrees too (points behind are divided by eps: huge, ill-conditioned)
    z_ok = np.abs(g["ref_cam"]).max(-1) < 50
    np.testing.assert_allclose(cam.cpu().numpy()[z_ok], g["ref_cam"][z_ok], rtol=1e-3, atol=1e-4)
