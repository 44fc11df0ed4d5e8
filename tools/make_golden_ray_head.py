"""Golden vectors for the head ray sampler, from the REFERENCE's own methods run on CPU here:
ViDARHeadBase._get_grid_features (vidar_head_base.py:420-509), the CE branch of loss (:586-592)
and the decode loop of get_point_cloud_prediction (:706-734, re-executed line by line on the
same tensors because the method needs dataset metas).  Writes tests/golden/ray_head.npz."""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import ray_cases as rc  # noqa: E402
from tools import ref_shim  # noqa: E402


def main():
    head = ref_shim.load("dense_heads.vidar_head_base").ViDARHeadBase
    c = rc.case()
    sig = [s.clone().requires_grad_(True) for s in c["sigma"]]
    self = types.SimpleNamespace(ray_grid_num=rc.NUM_WAY)
    lw = torch.tensor(rc.LOSS_W)
    r_mask, r_feat, r_w, r_len = head._get_grid_features(
        self, c["origin"], c["gt"], c["tindex"], sig, lw, rc.STEP)
    # CE branch of loss()
    r_label = r_mask.new_zeros(*r_feat.shape[:-1]).long()
    feat_t = r_feat.transpose(1, 2).contiguous()
    r_loss = F.cross_entropy(feat_t, r_label, reduction="none")
    loss = (r_loss * r_w).sum() / torch.clamp(r_w.sum(), min=1)
    loss.backward()
    rec = dict(r_mask=r_mask.numpy(), r_feat=r_feat.detach().numpy(), r_w=r_w.numpy(), r_len=r_len.numpy(),
               ce_per_ray=r_loss.detach().numpy(), loss=loss.detach().numpy(),
               grad_sigma0=sig[0].grad.numpy(), grad_sigma1=sig[1].grad.numpy())
    # decode loop (statements of get_point_cloud_prediction :700-732 on the same tensors)
    sigma = c["sigma"][-1]
    bs, Fr, Z, Y, X = sigma.shape
    pred = c["gt"].new_zeros(*c["gt"].shape[:2])
    idx_out = torch.full(c["gt"].shape[:2], -1, dtype=torch.long)
    r_grids = torch.from_numpy(np.arange(0, rc.NUM_WAY) + 0.5).to(c["gt"].dtype) * rc.STEP
    for b in range(bs):
        for f in range(Fr):
            cur_o = c["origin"][b, f:f + 1]
            sel = c["tindex"][b] == f
            cur_gt = c["gt"][b][sel]
            if len(cur_gt) == 0:
                continue
            cur_r = cur_gt - cur_o
            cur_rn = cur_r / torch.sqrt((cur_r ** 2).sum(-1, keepdims=True))
            grids = cur_o.view(-1, 1, 3) + cur_rn.view(-1, 1, 3) * r_grids.view(1, -1, 1)
            length = torch.sqrt(((grids - cur_o.view(-1, 1, 3)) ** 2).sum(-1))
            grids[..., 0] = grids[..., 0] / X
            grids[..., 1] = grids[..., 1] / Y
            grids[..., 2] = grids[..., 2] / Z
            grids = grids * 2 - 1
            cs = F.grid_sample(sigma[b, f].view(1, 1, Z, Y, X), grids.view(1, 1, *grids.shape))
            cs = cs.float().masked_fill((cs == 0), float("-inf")).squeeze(0).squeeze(0).squeeze(0)
            _, mi = cs.max(1)
            pred[b, sel] = torch.gather(length, dim=1, index=mi.view(-1, 1)).squeeze(-1)
            idx_out[b, sel] = mi
    rec.update(decode_pred=pred.numpy(), decode_idx=idx_out.numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ray_head.npz"), **rec)
    print("rays kept", r_feat.shape, "loss", float(loss))


if __name__ == "__main__":
    main()
