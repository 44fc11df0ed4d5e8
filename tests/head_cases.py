"""Seeded case for the ViDAR head loss / decode goldens (tools/make_golden_head.py and the tests)."""
import numpy as np
import torch

BEV_H, BEV_W, Z = 12, 16, 4
PC_RANGE = [-8.0, -6.0, -2.0, 8.0, 6.0, 2.0]          # 1 m voxels
FRAMES, INTER, BS = 2, 2, 2
RAY_GRID_NUM, RAY_GRID_STEP = 20, 1.0
LOSS_W = [[0.7], [1.3]]
HEAD_KW = dict(ray_grid_num=RAY_GRID_NUM, ray_grid_step=RAY_GRID_STEP, use_ce_loss=True, use_dist_loss=True,
               use_dense_loss=True, dense_loss_weight=0.5, eval_within_grid=False)
CALL_KW = dict(start_idx=0, tgt_bev_h=BEV_H, tgt_bev_w=BEV_W, tgt_pc_range=PC_RANGE)


def case(seed=0):
    g = torch.Generator().manual_seed(seed)
    preds = torch.randn(FRAMES, INTER, BS, BEV_H * BEV_W, Z, generator=g)
    lo = torch.tensor([-9.5, -7.0, -2.5])
    hi = torch.tensor([9.5, 7.0, 2.5])
    gt_points = []
    for n in (70, 55):
        xyz = lo + (hi - lo) * torch.rand(n, 3, generator=g)                   # a few outside the range
        extra = torch.rand(n, 1, generator=g)                                   # intensity column
        t = torch.randint(0, FRAMES + 1, (n, 1), generator=g).float()           # frame 2 is not predicted
        gt_points.append(torch.cat([xyz, extra, t], 1))
    origin = 0.4 * torch.randn(BS, FRAMES, 3, generator=g)
    return dict(pred_dict=dict(next_bev_preds=preds, valid_frames=[0, 1]), gt_points=gt_points, origin=origin,
                loss_weight=np.array(LOSS_W))
