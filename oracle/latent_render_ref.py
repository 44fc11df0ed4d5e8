"""CPU oracle for the LatentRendering core.  TEST INFRASTRUCTURE.

Restates projects/mmdet3d_plugin/bevformer/modules/ray_operations/latent_rendering.py:98-161
 with torch ops on channel-last maps: phase 1 = transmittance product to the
cell, phase 2 = probability-weighted pooling of the lora_a feature along the ray.  Pinned by
tests/golden/latent_rendering.npz, produced by running the reference LatentRendering class in
this container (tools/make_golden_latent.py)."""
import numpy as np
import torch
import torch.nn.functional as F


def _cells(Hb, Wb, dtype, device="cpu"):
    ys = torch.linspace(0.5, Hb - 0.5, Hb, dtype=dtype, device=device) / Hb
    xs = torch.linspace(0.5, Wb - 0.5, Wb, dtype=dtype, device=device) / Wb
    cy, cx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([cx.reshape(-1), cy.reshape(-1)], -1)          # [HW, 2] (x, y)


def latent_core(occ, feat, grid_num, grid_step, eps=1e-3, act="sigmoid", cells=None, prob_map=None):
    """occ [bs,Hb,Wb,D], feat [bs,Hb,Wb,D*G] -> prob [bs,Hb,Wb,D], pooled [bs,Hb*Wb,D*G].
    `cells` (a slice of the Hb*Wb cell indices) + `prob_map` restrict the work to a subset of
    cells for CPU timing samples (bench.py): outputs then cover only those cells."""
    bs, Hb, Wb, D = occ.shape
    Ca = feat.shape[-1]
    c = _cells(Hb, Wb, occ.dtype, occ.device)                                  # [HW,2]
    if cells is not None:
        c = c[cells]
    r = c - 0.5
    rn = torch.nan_to_num(r / torch.sqrt((r ** 2).sum(-1, keepdim=True)))
    t = torch.from_numpy(np.arange(grid_num) + 0.5).to(occ.dtype).to(occ.device) * (grid_step / (min(Hb, Wb) // 2))
    way = 0.5 + rn[:, None, :] * t[None, :, None]                     # [HW,G,2]
    grid = torch.cat([way, c[:, None, :]], 1) * 2 - 1                 # [HW,G+1,2]
    length = torch.sqrt((grid ** 2).sum(-1))                          # [HW,G+1]
    g_b = grid[None].expand(bs, -1, -1, -1)
    x = F.grid_sample(occ.permute(0, 3, 1, 2), g_b, align_corners=False)     # [bs,D,HW,G+1]
    a = torch.sigmoid(x) if act == "sigmoid" else 1 - torch.exp(-F.relu(x))
    m = (length < length[:, -1:]).to(occ.dtype)                       # [HW,G+1]
    trans = torch.prod(1 - a * m[None, None], -1)                     # [bs,D,HW]
    prob = (trans * a[..., -1]).permute(0, 2, 1)
    if cells is None:
        prob = prob.reshape(bs, Hb, Wb, D)
        prob_map = prob
    # phase 2
    gw = g_b[:, :, :-1]
    fs = F.grid_sample(feat.permute(0, 3, 1, 2), gw, align_corners=False)    # [bs,Ca,HW,G]
    ps = F.grid_sample(prob_map.permute(0, 3, 1, 2), gw, align_corners=False)    # [bs,D,HW,G]
    bound = torch.minimum(1 / rn[:, 0].abs(), 1 / rn[:, 1].abs())     # [HW]
    ps = ps * (length[:, :-1] < bound[:, None]).to(occ.dtype)[None, None]
    pi = ps / (ps.sum(-1, keepdim=True) + eps)
    n_cells = c.shape[0]
    pooled = (fs.view(bs, D, Ca // D, n_cells, grid_num) * pi[:, :, None]).sum(-1)   # [bs,D,g,HW]
    pooled = pooled.reshape(bs, Ca, n_cells).permute(0, 2, 1)
    return prob, pooled
