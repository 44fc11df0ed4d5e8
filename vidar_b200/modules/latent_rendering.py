"""LatentRendering module on the fused CUDA core (vidar_b200/csrc/latent_render.cu).

Same registry name, constructor kwargs, parameter names (`unsup_raymarching_head.*`,
`lora_a.*`, `lora_b.*`), call signature and numerics as
projects/mmdet3d_plugin/bevformer/modules/ray_operations/latent_rendering.py:37-162; it is
instantiated from the `latent_render=dict(...)` kwarg of the encoder/decoder layers
(encoder_v2.py:81-82) and called as `latent_render(query.view(bs, bev_h, bev_w, C))`.

The three Linear layers and the final product stay PyTorch ops (cuBLAS); the ray-marching
between them -- grid_sample x3, masks, cumprod, normalisation, pooling in the reference -- is
one custom autograd op (two kernels forward, two backward) that never materialises the
[bs, 16, 40000, 257] tensors.  CUDA only.
"""
import torch
import torch.nn as nn

from .. import _lib
from ..registry import ATTENTION, BaseModule

_ACT = {"exp": 0, "sigmoid": 1}


class _LatentRenderCore(torch.autograd.Function):
    """(occ [bs,Hb,Wb,D], feat [bs,Hb,Wb,D*G]) -> (prob [bs,Hb,Wb,D], pooled [bs,Hb*Wb,D*G])."""

    @staticmethod
    def forward(ctx, occ, feat, grid_num, grid_step, eps, act):
        _lib.require_cuda(occ=occ.contiguous(), feat=feat.contiguous())
        occ, feat = occ.float().contiguous(), feat.float().contiguous()
        bs, Hb, Wb, D = occ.shape
        Ca = feat.shape[-1]
        if feat.shape[:3] != occ.shape[:3] or Ca % D != 0:
            raise RuntimeError("feat must be [bs, Hb, Wb, pred_height * k]")
        prob = torch.empty_like(occ)
        pooled = torch.empty((bs, Hb * Wb, Ca), dtype=torch.float32, device=occ.device)
        with torch.cuda.device(occ.device):
            _lib.check(_lib.lib().vidar_latent_render_forward(
                _lib.ptr(occ), _lib.ptr(feat), _lib.ptr(prob), _lib.ptr(pooled), bs, D, Ca // D, Hb, Wb,
                int(grid_num), float(grid_step), float(eps), int(act), _lib.stream_ptr(occ.device)))
        ctx.save_for_backward(occ, feat, prob)
        ctx.cfg = (int(grid_num), float(grid_step), float(eps), int(act))
        return prob, pooled

    @staticmethod
    def backward(ctx, grad_prob, grad_pooled):
        occ, feat, prob = ctx.saved_tensors
        grid_num, grid_step, eps, act = ctx.cfg
        bs, Hb, Wb, D = occ.shape
        Ca = feat.shape[-1]
        grad_prob = grad_prob.float().contiguous()
        grad_pooled = grad_pooled.float().contiguous()
        scratch = torch.empty_like(occ)
        grad_occ = torch.zeros_like(occ)
        grad_feat = torch.zeros_like(feat)
        with torch.cuda.device(occ.device):
            _lib.check(_lib.lib().vidar_latent_render_backward(
                _lib.ptr(occ), _lib.ptr(feat), _lib.ptr(prob), _lib.ptr(grad_prob), _lib.ptr(grad_pooled),
                _lib.ptr(scratch), _lib.ptr(grad_occ), _lib.ptr(grad_feat), bs, D, Ca // D, Hb, Wb,
                grid_num, grid_step, eps, act, _lib.stream_ptr(occ.device)))
        return grad_occ, grad_feat, None, None, None, None


latent_render_core = _LatentRenderCore.apply


@ATTENTION.register_module()
class LatentRendering(BaseModule):
    """Ray-marching adaptor ("latent rendering") of ViDAR, latent_rendering.py:37-162."""

    def __init__(self, embed_dims=256, num_pred_fcs=2, pred_height=1, grid_num=128, grid_step=0.5,
                 reduction=16, act="exp", viz_response=False, init_cfg=None):
        super().__init__(init_cfg)
        self.embed_dims = embed_dims
        self.num_pred_fcs = num_pred_fcs
        self.grid_num = grid_num
        self.grid_step = grid_step
        self.viz_response = viz_response
        self.act = act
        branch = []
        for _ in range(self.num_pred_fcs):
            branch.append(nn.Linear(self.embed_dims, self.embed_dims))
            branch.append(nn.LayerNorm(self.embed_dims))
            branch.append(nn.ReLU(inplace=True))
        branch.append(nn.Linear(self.embed_dims, pred_height))
        self.unsup_raymarching_head = nn.Sequential(*branch)
        self.pred_height = pred_height
        self.lora_a = nn.Linear(self.embed_dims, self.embed_dims // reduction)
        self.lora_b = nn.Linear(self.embed_dims // reduction, self.embed_dims)

    def forward(self, embed, eps=1e-3, **kwargs):
        """embed [bs, bev_h, bev_w, embed_dims] -> same shape."""
        if self.act not in _ACT:
            raise NotImplementedError("Only support exp or sigmoid activation_fn for now.")
        bs, bev_h, bev_w, _ = embed.shape
        occ = self.unsup_raymarching_head(embed)            # [bs, h, w, pred_height]   (:94)
        feat = self.lora_a(embed)                           # [bs, h, w, embed/reduction] (:134)
        prob, pooled = latent_render_core(occ, feat, self.grid_num, self.grid_step, eps, _ACT[self.act])
        out = self.lora_b(pooled).view(bs, bev_h, bev_w, self.embed_dims)       # (:153-155)
        out = out.view(bs, bev_h, bev_w, self.pred_height, -1) * prob.view(bs, bev_h, bev_w, self.pred_height, 1)
        return out.view(bs, bev_h, bev_w, self.embed_dims)
