"""CPU oracle for the ViDAR head ray sampler / CE loss / arg-max decode.  TEST INFRASTRUCTURE.

Restates projects/mmdet3d_plugin/bevformer/dense_heads/vidar_head_base.py
  _get_grid_features :420-509, the CE branch of loss :586-592, and the decode loop of
  get_point_cloud_prediction :700-734
with plain torch ops (F.grid_sample for the trilinear interpolation, as the reference does).
Pinned by tests/golden/ray_head.npz, produced by running the reference's own methods in this
container (tools/make_golden_ray_head.py).
"""
import torch
import torch.nn.functional as F


def sample_frame(sigma_f, origin, gt, num_way, step, with_gt=True):
    """One frame.  sigma_f [Z,Y,X]; origin [3]; gt [R,3] (voxel units).
    -> logits [R,K] (-inf where masked), length [R,K], valid [R] bool."""
    Z, Y, X = sigma_f.shape
    o = origin.view(1, 1, 3)
    ray = gt - origin.view(1, 3)
    unit = ray / torch.sqrt((ray ** 2).sum(-1, keepdim=True))
    t = (torch.arange(num_way, dtype=torch.float64) + 0.5).to(gt.dtype) * step
    pts = o + unit.unsqueeze(1) * t.view(1, -1, 1)
    if with_gt:
        pts = torch.cat([gt.unsqueeze(1), pts], 1)
    length = torch.sqrt(((pts - o) ** 2).sum(-1))
    size = torch.tensor([X, Y, Z], dtype=gt.dtype)
    norm = pts / size * 2 - 1
    masked = ((norm <= -1.) | (norm >= 1)).any(-1)
    valid = ((norm[:, 0] > -1.) & (norm[:, 0] < 1.)).all(-1) if with_gt else torch.ones(len(gt), dtype=torch.bool)
    val = F.grid_sample(sigma_f.view(1, 1, Z, Y, X), norm.view(1, 1, *norm.shape), mode="bilinear",
                        padding_mode="zeros", align_corners=False).view(norm.shape[:2])
    logits = val + torch.zeros_like(val).masked_fill(masked, float("-inf"))
    return logits, length, valid


def grid_features(origin_grids, gt_grids, gt_tindex, intermediate_sigma, loss_weights, step, num_way):
    """All four outputs of _get_grid_features for lists over (batch, frame)."""
    bs, Fr = intermediate_sigma[0].shape[:2]
    masks, feats, weights, lengths = [], [], [], []
    for b in range(bs):
        for f in range(Fr):
            gt = gt_grids[b][gt_tindex[b] == f]
            per_lvl, keep, length = [], None, None
            for sig in intermediate_sigma:
                logits, length, keep = sample_frame(sig[b, f], origin_grids[b, f], gt, num_way, step)
                per_lvl.append(logits[keep])
            feats.append(torch.stack(per_lvl, 0))
            lengths.append(length[keep])
            masks.append(torch.isinf(per_lvl[0]))
            w = torch.stack([torch.full((int(keep.sum()),), float(loss_weights[f][l]))
                             for l in range(len(intermediate_sigma))], 0)
            weights.append(w)
    mask = torch.cat(masks, 0)
    r_mask = torch.zeros(mask.shape).masked_fill(mask, float("-inf"))
    return r_mask, torch.cat(feats, 1), torch.cat(weights, 1), torch.cat(lengths, 0)


def ce_loss(r_feat_total, r_loss_weight_total):
    """loss :586-592 -- cross entropy with label 0 over the K logits of every ray."""
    logp0 = torch.log_softmax(r_feat_total, -1)[..., 0]
    return (-(logp0) * r_loss_weight_total).sum() / torch.clamp(r_loss_weight_total.sum(), min=1)


def decode_frame(sigma_f, origin, gt, num_way, step):
    """get_point_cloud_prediction :712-732 for one frame -> (pred_dist [R], argmax index [R])."""
    logits, length, _ = sample_frame(sigma_f, origin, gt, num_way, step, with_gt=False)
    Z, Y, X = sigma_f.shape
    # the decode path applies no boundary mask: outside samples are exact zeros -> -inf
    o = origin.view(1, 1, 3)
    ray = gt - origin.view(1, 3)
    unit = ray / torch.sqrt((ray ** 2).sum(-1, keepdim=True))
    t = (torch.arange(num_way, dtype=torch.float64) + 0.5).to(gt.dtype) * step
    pts = o + unit.unsqueeze(1) * t.view(1, -1, 1)
    norm = pts / torch.tensor([X, Y, Z], dtype=gt.dtype) * 2 - 1
    val = F.grid_sample(sigma_f.view(1, 1, Z, Y, X), norm.view(1, 1, *norm.shape), mode="bilinear",
                        padding_mode="zeros", align_corners=False).view(norm.shape[:2])
    val = val.masked_fill(val == 0, float("-inf"))
    idx = val.max(1)[1]
    return torch.gather(length, 1, idx.view(-1, 1)).squeeze(-1), idx
