"""Host logic of the ViDAR head mirror (vidar_b200/head.py) that needs no GPU, against goldens from
the REFERENCE methods (tools/make_golden_head.py)."""
import os

import numpy as np
import torch

from tests import head_cases as hc
from vidar_b200.head import ViDARRayHead, get_inside_mask

GOLD = os.path.join(os.path.dirname(__file__), "golden", "head.npz")


def _head():
    c = hc.case()
    return ViDARRayHead(loss_weight=c["loss_weight"], **hc.HEAD_KW), c


def test_process_gt_points_matches_reference():
    g = np.load(GOLD)
    head, c = _head()
    preds = c["pred_dict"]["next_bev_preds"][:, -1:]
    og, op, gg, gp, gt = head._process_gt_points(preds, c["gt_points"], c["origin"], [0, 1], 0, hc.FRAMES, hc.BEV_H,
                                                 hc.BEV_W, hc.PC_RANGE)
    np.testing.assert_array_equal(og.numpy(), g["origin_grids"])
    # same rays in the same order as the reference; the padded length is the input length here (the reference
    # drops other frames' points before padding -- a host sync), so compare the reference's prefix and
    # require pure padding behind it
    k = g["gt_grids"].shape[1]
    assert gg.shape[1] == max(len(p) for p in c["gt_points"]) >= k
    np.testing.assert_array_equal(gg.numpy()[:, :k], g["gt_grids"])   # NaN padding compares equal
    np.testing.assert_array_equal(gp.numpy()[:, :k], g["gt_points"])
    np.testing.assert_array_equal(gt.numpy()[:, :k], g["gt_tindex"])
    assert np.isnan(gg.numpy()[:, k:]).all() and (gt.numpy()[:, k:] == -1).all()
    assert (gt == -1).any() and gt.max() == hc.FRAMES - 1
    # origin defaults to the ego position (zeros) when not given
    og0, op0, *_ = head._process_gt_points(preds, c["gt_points"], None, [0, 1], 0, hc.FRAMES, hc.BEV_H, hc.BEV_W, hc.PC_RANGE)
    assert float(op0.abs().sum()) == 0 and torch.allclose(og0[0, 0], torch.tensor([hc.BEV_W / 2, hc.BEV_H / 2, hc.Z / 2]))


def _reference_gumbel_statements(grid_embed, grid_length):
    """vidar_head_base.py:754-773, statement by statement (test-side restatement)."""
    import torch.nn.functional as F
    pred_dist = F.gumbel_softmax(grid_embed, hard=True)
    pred_dist = (pred_dist * grid_length).sum(-1).detach()
    grid_embed = grid_embed - grid_embed.max(-1, keepdims=True)[0]
    exp_embed = torch.exp(grid_embed)
    exp_whole = exp_embed.sum(-1)
    next_ind = (grid_length > pred_dist.unsqueeze(-1)).float()
    prob_next = (exp_embed * next_ind).sum(-1) / exp_whole
    prob_next = 1 - prob_next.detach() + prob_next
    return prob_next * pred_dist


def test_gumbel_distance_equals_reference_statements_and_given_noise():
    head, _ = _head()
    g = torch.Generator().manual_seed(3)
    embed = torch.randn(2, 17, 20, generator=g)
    embed[0, 3, 12:] = float("-inf")                      # waypoints outside the volume
    length = (torch.arange(20.0) + 0.5).expand(2, 17, 20)
    e1, e2, e3 = (embed.clone().requires_grad_(True) for _ in range(3))
    torch.manual_seed(5)
    ref = _reference_gumbel_statements(e1, length)
    torch.manual_seed(5)
    a = head._custom_gumbel_softmax_distance(e2, length)  # draws its own noise: same generator consumption
    torch.manual_seed(5)
    noise = -torch.empty_like(embed).exponential_().log()
    b = head._custom_gumbel_softmax_distance(e3, length, noise)
    assert torch.equal(a, ref) and torch.equal(b, ref)
    assert set(np.unique(ref.detach().numpy())) <= set(np.unique(length.numpy()))
    w = torch.randn(2, 17, generator=g)
    for out, e in ((ref, e1), (a, e2), (b, e3)):
        (out * w).sum().backward()
    assert torch.allclose(e2.grad, e1.grad, rtol=1e-6, atol=1e-7) and torch.equal(e2.grad, e3.grad)
    assert float(e1.grad.abs().sum()) > 0 and float(e1.grad[0, 3, 12:].abs().sum()) == 0


def test_rendered_pcds_and_inside_mask():
    head, _ = _head()
    origin = torch.tensor([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]])
    pts = torch.tensor([[[2.0, 0.0, 0.0], [0.0, 3.0, 0.0], [1.0, 0.0, 4.0], [9.0, 9.0, 9.0]]])
    tindex = torch.tensor([[0.0, 0.0, 1.0, -1.0]])
    gt_dist = torch.tensor([[2.0, 0.0, 4.0, 1.0]])                   # second ray masked by gt_dist == 0
    pred = torch.tensor([[1.0, 5.0, 2.0, 7.0]])
    pcds = head.get_rendered_pcds(origin, pts, tindex, gt_dist, pred, hc.PC_RANGE)
    assert torch.allclose(pcds[0][0], torch.tensor([[1.0, 0.0, 0.0]]))
    assert torch.allclose(pcds[0][1], torch.tensor([[1.0, 0.0, 2.0]]))
    head.eval_within_grid = True
    pcds = head.get_rendered_pcds(origin, pts, tindex, gt_dist, pred, [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    assert pcds[0][0].shape[0] == 0 and pcds[0][1].shape[0] == 0     # both GT points lie outside that box
    assert get_inside_mask(pts[0], hc.PC_RANGE).tolist() == [True, True, False, False]
