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
        np.testing.assert_allclose(float(v.detach()), float(g[f"loss_{k}"]), rtol=1e-4, err_msg=k)
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


def test_dense_term_fused_decode_equals_torch_statements(cuda):
    """loss.dense_voxel through the fused sampler + gumbel kernel (default) and through the
    materialising sampler + the reference's torch statements: same value, same gradient."""
    g = np.load(GOLD)
    res = []
    for fused in (True, False):
        head, c, gt = _setup(cuda)
        head.use_ce_loss = head.use_dist_loss = False
        head.fuse_dense_decode = fused
        preds = c["pred_dict"]["next_bev_preds"].to(cuda).requires_grad_(True)
        noise = {"dense": torch.from_numpy(g["gumbel_dense"]).to(cuda)}
        losses = head.loss(dict(next_bev_preds=preds, valid_frames=[0, 1]), gt, pred_frame_num=hc.FRAMES,
                           batched_origin_points=c["origin"].to(cuda), gumbels=noise, **hc.CALL_KW)
        assert set(losses) == {"loss.dense_voxel"}
        losses["loss.dense_voxel"].backward()
        res.append((float(losses["loss.dense_voxel"]), preds.grad.clone()))
        np.testing.assert_allclose(res[-1][0], float(g["loss_loss.dense_voxel"]), rtol=1e-4)
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0])
    err = (res[0][1] - res[1][1]).abs().max().item()
    assert err <= 1e-5 * res[1][1].abs().max().item(), f"grad: {err:.3e}"


def test_gumbel_distance_kernel_vs_torch_statements(cuda):
    """ray_head.gumbel_distance at a larger size: rays that leave the volume (masked waypoints),
    rays whose end point is outside (dist 0, no gradient), every frame."""
    from vidar_b200 import ray_head
    gen = torch.Generator().manual_seed(3)
    F_, Z, Y, X, R, K = 3, 8, 40, 36, 3000, 64
    sigma = torch.randn(F_, Z, Y, X, generator=gen).to(cuda)
    origin = (torch.tensor([X / 2, Y / 2, Z / 2]) + torch.randn(F_, 3, generator=gen)).to(cuda)
    pts = (torch.rand(R, 3, generator=gen) * torch.tensor([X + 4.0, Y + 4.0, Z + 2.0]) - torch.tensor([2.0, 2.0, 1.0])).to(cuda)
    frame = torch.randint(0, F_, (R,), generator=gen).sort()[0].to(torch.int32).to(cuda)
    noise = (-torch.empty(R, K).exponential_(generator=gen).log()).to(cuda)
    gout = torch.randn(R, generator=gen).to(cuda)

    s1 = sigma.clone().requires_grad_(True)
    d1 = ray_head.gumbel_distance(s1, origin, pts, frame, K, 1.0, noise)
    (d1 * gout).sum().backward()

    s2 = sigma.clone().requires_grad_(True)
    logits, length, valid = ray_head.ray_sample(s2, origin, pts, frame, K, 1.0, True)
    head = ViDARRayHead(loss_weight=np.array([[1.0]]), ray_grid_num=K)
    keep = valid > 0
    d2 = torch.zeros(R, device=cuda)
    d2[keep] = head._custom_gumbel_softmax_distance(logits[keep][:, 1:], length[keep][:, 1:], noise[keep])
    (d2 * gout).sum().backward()

    assert 0.05 < float((~keep).float().mean()) < 0.6 and bool(torch.isinf(logits[keep]).any())
    assert float(d1[~keep].abs().max()) == 0
    assert torch.equal(d1, d2.detach())                     # the decoded length is the waypoint's length, bit for bit
    err = (s1.grad - s2.grad).abs().max().item()
    assert err <= 2e-5 * s2.grad.abs().max().item(), f"grad_sigma: {err:.3e} vs {s2.grad.abs().max().item():.3e}"
