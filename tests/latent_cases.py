"""Seeded module-level case for LatentRendering goldens/tests."""
import torch

CFG = dict(type="LatentRendering", embed_dims=64, num_pred_fcs=0, pred_height=16, grid_num=48,
           grid_step=0.5, reduction=4, act="sigmoid")
CFG_EXP = dict(type="LatentRendering", embed_dims=64, num_pred_fcs=1, pred_height=4, grid_num=40,
               grid_step=0.5, reduction=8, act="exp")      # 8 channels / 4 heights: G = 2
CFG_D1 = dict(type="LatentRendering", embed_dims=64, num_pred_fcs=0, pred_height=1, grid_num=40,
              grid_step=0.5, reduction=4, act="sigmoid")    # the class defaults' shape: 1 height, 16 channels
BEV = (20, 24)


def seeded_state(module, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, p in sorted(module.state_dict().items()):
        scale = 0.5 if name.endswith("bias") else 1.5 / (p.shape[-1] ** 0.5)
        sd[name] = torch.randn(p.shape, generator=g) * scale
    return sd


def case(seed=0, bs=2):
    g = torch.Generator().manual_seed(seed)
    return dict(embed=torch.randn(bs, BEV[0], BEV[1], 64, generator=g),
                grad=torch.randn(bs, BEV[0], BEV[1], 64, generator=g))
