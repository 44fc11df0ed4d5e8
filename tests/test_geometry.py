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


def test_value_from_fpn_equals_flatten_then_project():
    """value built straight from the FPN maps (embeddings folded into the bias) equals the
    reference's flatten + embeds (transformer.py:159-179) followed by SCA's permute/reshape and
    MSDeformableAttention3D's value_proj (spatial_cross_attention.py:158-160,333-336)."""
    import torch
    from vidar_b200 import bev_geometry as bg
    g = torch.Generator().manual_seed(0)
    bs, cams, C, H = 2, 3, 32, 4
    feats = [torch.randn(bs, cams, C, h, w, generator=g) for h, w in ((6, 10), (3, 5), (2, 3))]
    cam_e, lvl_e = torch.randn(cams, C, generator=g), torch.randn(4, C, generator=g)
    proj = torch.nn.Linear(C, C)
    flat, shapes, lsi = bg.flatten_fpn(feats, cam_e, lvl_e)
    assert flat.shape == (cams, 60 + 15 + 6, bs, C) and shapes.tolist() == [[6, 10], [3, 5], [2, 3]] and lsi.tolist() == [0, 60, 75]
    # the reference's own statements on one element: level 1, camera 2, batch 1, pixel (2, 3)
    ref = feats[1][1, 2, :, 2, 3] + cam_e[2] + lvl_e[1]
    assert torch.allclose(flat[2, 60 + 2 * 5 + 3, 1], ref)
    want = proj(flat.permute(2, 0, 1, 3).reshape(bs * cams, -1, C)).view(bs * cams, -1, H, C // H)
    got, s2, l2 = bg.value_from_fpn(feats, cam_e, lvl_e, proj, H)
    assert torch.equal(s2, shapes) and torch.equal(l2, lsi)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    got2, _, _ = bg.value_from_fpn(feats, None, lvl_e, proj, H)
    flat2, _, _ = bg.flatten_fpn(feats, None, lvl_e)
    want2 = proj(flat2.permute(2, 0, 1, 3).reshape(bs * cams, -1, C)).view(bs * cams, -1, H, C // H)
    assert torch.allclose(got2, want2, rtol=1e-5, atol=1e-5)
