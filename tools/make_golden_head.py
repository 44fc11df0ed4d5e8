"""Golden vectors for the head's loss / decode logic from the REFERENCE methods run on CPU here:
ViDARHeadBase.loss (all three terms), get_point_cloud_prediction, _process_gt_points,
get_rendered_pcds, _custom_gumbel_softmax_distance (vidar_head_base.py:219-276, 344-389, 510-773).

    python tools/make_golden_head.py          # writes tests/golden/head.npz

Two things are supplied from outside the reference tree:
  * `chamfer_distance` -- mmdet3d 0.17.1 is not installed; restated below from its published
    definition (criterion 'l2' = squared distance summed over xyz, min over the other cloud,
    reduction 'mean');
  * the helpers of utils/e2e_predictor_utils.py, executed function-by-function from the reference
    file (its module body JIT-compiles CUDA), see tools/ref_shim.py::install_e2e_utils.
The Gumbel noise the reference draws (CPU generator, seed 11) is re-drawn here in the same order and
stored, so the GPU test feeds the identical noise.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import head_cases as hc  # noqa: E402
from tools import ref_shim  # noqa: E402

SEED = 11


def chamfer_distance(src, dst, src_weight=1.0, dst_weight=1.0, criterion_mode="l2", reduction="mean"):
    assert criterion_mode == "l2" and reduction == "mean"
    d = ((src.unsqueeze(2) - dst.unsqueeze(1)) ** 2).sum(-1)          # [B, N, M]
    s2d, i1 = torch.min(d, dim=2)
    d2s, i2 = torch.min(d, dim=1)
    return torch.mean(s2d * src_weight), torch.mean(d2s * dst_weight), i1, i2


def main():
    utils = ref_shim.install_e2e_utils()
    f3 = utils.get_bev_grids_3d
    utils.get_bev_grids_3d = lambda H, W, Z, bs=1, device="cpu", dtype=torch.float: f3(H, W, Z, bs, "cpu", dtype)
    mod = ref_shim.load("dense_heads.vidar_head_base")
    mod.chamfer_distance = chamfer_distance
    mod.e2e_predictor_utils = utils
    head = object.__new__(mod.ViDARHeadBase)
    torch.nn.Module.__init__(head)
    for k, v in hc.HEAD_KW.items():
        setattr(head, k, v)
    c = hc.case()
    head.loss_weight = c["loss_weight"]
    preds = c["pred_dict"]["next_bev_preds"].clone().requires_grad_(True)
    pd = dict(next_bev_preds=preds, valid_frames=c["pred_dict"]["valid_frames"])
    shapes = []
    orig = head._custom_gumbel_softmax_distance
    head._custom_gumbel_softmax_distance = lambda e, l: (shapes.append(tuple(e.shape)), orig(e, l))[1]
    torch.manual_seed(SEED)
    losses = head.loss(pd, c["gt_points"], pred_frame_num=hc.FRAMES, batched_origin_points=c["origin"], **hc.CALL_KW)
    total = sum(losses.values())
    total.backward()
    rec = {f"loss_{k}": v.detach().numpy() for k, v in losses.items()}
    rec["grad_preds"] = preds.grad.numpy()
    # the reference drew its Gumbel noise in this order: dist term, then dense term
    torch.manual_seed(SEED)
    for name, shp in zip(("dist", "dense"), shapes):
        rec[f"gumbel_{name}"] = (-torch.empty(shp).exponential_().log()).numpy()
    print({k: float(v) for k, v in losses.items()}, shapes)
    # ground-truth preparation
    og, op, gg, gp, gt = head._process_gt_points(preds.detach()[:, -1:], c["gt_points"], c["origin"], [0, 1], 0, hc.FRAMES,
                                                 hc.BEV_H, hc.BEV_W, hc.PC_RANGE)
    rec.update(origin_grids=og.numpy(), gt_grids=gg.numpy(), gt_points=gp.numpy(), gt_tindex=gt.numpy())
    # decode
    with torch.no_grad():
        dec = head.get_point_cloud_prediction(dict(next_bev_preds=preds.detach(), valid_frames=[0, 1]), c["gt_points"],
                                              batched_origin_points=c["origin"], **hc.CALL_KW)
    for key in ("pred_pcds", "gt_pcds"):
        for b in range(hc.BS):
            for t in range(hc.FRAMES):
                rec[f"{key}_{b}_{t}"] = dec[key][b][t].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "head.npz"), **rec)


if __name__ == "__main__":
    main()
