"""ViDAR head loss / decode on the GPU (CUDA sampler, fused CE, NN kernel) against goldens from the
REFERENCE methods ViDARHeadBase.loss / get_point_cloud_prediction (tools/make_golden_head.py)."""
import os

import numpy as np
import pytest
import torch

from tests import head_cases as hc
from vidar_b200.head import ViDARRayHead

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "head.npz")


def _setup(cuda):
    c = hc.case()
    head = ViDARRayHead(loss_weight=c["loss_weight"], **hc.HEAD_KW)
    gt = [p.to(cuda) for p in c["gt_points"]]
    return head, c, gt


def test_loss_terms_and_gradient_match_reference(cuda):
    g = np.load(GOLD)
    head, c, gt = _setup(cuda)
    preds = c["pred_dict"]["next_bev_preds"].to(cuda).requires_grad_(True)
    noise = {k: torch.from_numpy(g[f"gumbel_{k}"]).to(cuda) for k in ("dist", "dense")}
    losses = head.loss(dict(next_bev_preds=preds, valid_frames=[0, 1]), gt, pred_frame_num=hc.FRAMES,
                       batched_origin_points=c["origin"].to(cuda), gumbels=noise, **hc.CALL_KW)
    assert set(losses) == {"dist.loss", "regularization.loss", "loss.dense_voxel"}
    for k, v in losses.items():
        np.testing.assert_allclose(float(v), float(g[f"loss_{k}"]), rtol=1e-4, err_msg=k)
    sum(losses.values()).backward()
    ref = g["grad_preds"]
    err = np.abs(preds.grad.cpu().numpy() - ref).max()
    assert err <= 1e-4 * np.abs(ref).max(), f"grad wrt next_bev_preds: max err {err:.3e} vs {np.abs(ref).max():.3e}"
    assert float(preds.grad[:, 0].abs().sum()) == 0          # only the last intermediate output is supervised


def test_ce_only_configuration_and_default_noise(cuda):
    """The shipped loss configuration draws its own Gumbel noise; CE must not depend on it."""
    g = np.load(GOLD)
    head, c, gt = _setup(cuda)
    head.use_dist_loss = False
    preds = c["pred_dict"]["next_bev_preds"].to(cuda)
    losses = head.loss(dict(next_bev_preds=preds, valid_frames=[0, 1]), gt, pred_frame_num=hc.FRAMES,
                       batched_origin_points=c["origin"].to(cuda), **hc.CALL_KW)
    assert set(losses) == {"regularization.loss", "loss.dense_voxel"}
    np.testing.assert_allclose(float(losses["regularization.loss"]), float(g["loss_regularization.loss"]), rtol=1e-4)
    assert np.isfinite(float(losses["loss.dense_voxel"])) and float(losses["loss.dense_voxel"]) > 0


def test_point_cloud_prediction_matches_reference(cuda):
    g = np.load(GOLD)
    head, c, gt = _setup(cuda)
    preds = c["pred_dict"]["next_bev_preds"].to(cuda)
    dec = head.get_point_cloud_prediction(dict(next_bev_preds=preds, valid_frames=[0, 1]), gt,
                                          batched_origin_points=c["origin"].to(cuda), **hc.CALL_KW)
    assert torch.equal(dec["origin"].cpu(), c["origin"])
    n = 0
    for key in ("pred_pcds", "gt_pcds"):
        for b in range(hc.BS):
            for t in range(hc.FRAMES):
                ref = g[f"{key}_{b}_{t}"]
                out = dec[key][b][t].cpu().numpy()
                assert out.shape == ref.shape
                np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)
                n += ref.shape[0]
    assert n > 100
