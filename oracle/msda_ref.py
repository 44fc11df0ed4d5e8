"""CPU oracle for multi-scale deformable attention.  TEST INFRASTRUCTURE ONLY.

The arithmetic of this path lives in a dependency that is NOT vendored in the reference:
mmcv-full==1.4.0 (README.md:103 of the reference), `_ext.ms_deform_attn_{forward,backward}`
and its any-device twin `mmcv.ops.multi_scale_deform_attn.multi_scale_deformable_attn_pytorch`
(imported at projects/mmdet3d_plugin/bevformer/modules/spatial_cross_attention.py:8 and used on
non-CUDA tensors at :393-394).  This file restates the published algorithm (Deformable-DETR,
Zhu et al. 2020, as implemented by mmcv): per level, bilinear `grid_sample(padding='zeros',
align_corners=False)` of the value map at `2*loc-1`, weighted by the attention weights and
summed over levels x points.

Two independent statements are given:
  * `msda_grid_sample`  -- the grid_sample formulation (what the reference's CPU path runs);
    autograd provides its backward;
  * `msda_explicit`     -- the per-corner formulation the CUDA kernels follow
    (pixel = loc*size - 0.5, four individually zero-padded corners,
    analytic gradients), in numpy float64.
Pinning: tests/test_oracle_msda.py checks both against each other and against the
same-lineage implementation shipped in `transformers`
(`transformers.models.deformable_detr.modeling_deformable_detr`), which is present in this
image; the reference itself holds no golden vectors for this op.
"""
import numpy as np
import torch
import torch.nn.functional as F


def msda_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """value [B,K,H,C]; spatial_shapes [L,2] (h,w); sampling_locations [B,Q,H,L,P,2] (x,y in
    [0,1]); attention_weights [B,Q,H,L,P] -> [B,Q,H*C].  Differentiable torch ops only."""
    B, K, H, C = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in torch.as_tensor(spatial_shapes).tolist()]
    assert sum(h * w for h, w in shapes) == K
    out = value.new_zeros(B * H, C, Q)
    start = 0
    # [B,Q,H,L,P] -> [B*H, Q, L, P]
    w_all = attention_weights.permute(0, 2, 1, 3, 4).reshape(B * H, Q, L, P)
    for lvl, (h, w) in enumerate(shapes):
        # value of this level as an image batch [B*H, C, h, w]
        v = value[:, start:start + h * w].permute(0, 2, 3, 1).reshape(B * H, C, h, w)
        start += h * w
        # grid [B*H, Q, P, 2] in [-1, 1]
        g = sampling_locations[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(B * H, Q, P, 2)
        s = F.grid_sample(v, 2.0 * g - 1.0, mode="bilinear", padding_mode="zeros",
                          align_corners=False)            # [B*H, C, Q, P]
        out = out + (s * w_all[:, None, :, lvl, :]).sum(-1)
    return out.reshape(B, H * C, Q).transpose(1, 2).contiguous()


def msda_grid_sample_backward(value, spatial_shapes, sampling_locations, attention_weights,
                              grad_output, dtype=torch.float64):
    """(grad_value, grad_loc, grad_attn) of msda_grid_sample by autograd, in `dtype`."""
    v = value.detach().to(dtype).requires_grad_(True)
    loc = sampling_locations.detach().to(dtype).requires_grad_(True)
    aw = attention_weights.detach().to(dtype).requires_grad_(True)
    out = msda_grid_sample(v, spatial_shapes, loc, aw)
    out.backward(grad_output.to(dtype))
    return v.grad, loc.grad, aw.grad


def msda_explicit(value, spatial_shapes, level_start_index, sampling_locations,
                  attention_weights, grad_output=None):
    """Per-corner statement in numpy float64.  Returns out, or (out, gv, gloc, gattn) when
    grad_output is given.  Python loops over (level, point) only; vectorised over b,q,h."""
    value = np.asarray(value, np.float64)
    loc = np.asarray(sampling_locations, np.float64)
    aw = np.asarray(attention_weights, np.float64)
    shapes = np.asarray(spatial_shapes).astype(np.int64)
    lsi = np.asarray(level_start_index).astype(np.int64)
    B, K, H, C = value.shape
    _, Q, _, L, P, _ = loc.shape
    out = np.zeros((B, Q, H, C))
    if grad_output is not None:
        go = np.asarray(grad_output, np.float64).reshape(B, Q, H, C)
        gv = np.zeros_like(value)
        gloc = np.zeros_like(loc)
        gaw = np.zeros_like(aw)
    bi = np.arange(B)[:, None, None]
    hi = np.arange(H)[None, None, :]
    for l in range(L):
        Hl, Wl = int(shapes[l, 0]), int(shapes[l, 1])
        for p in range(P):
            x = loc[:, :, :, l, p, 0] * Wl - 0.5
            y = loc[:, :, :, l, p, 1] * Hl - 0.5
            ok = (y > -1) & (x > -1) & (y < Hl) & (x < Wl)
            xs = np.where(ok, x, 0.0)
            ys = np.where(ok, y, 0.0)
            x0 = np.floor(xs).astype(np.int64)
            y0 = np.floor(ys).astype(np.int64)
            lw = xs - x0
            lh = ys - y0
            hw, hh = 1 - lw, 1 - lh
            corners = []
            for (yy, xx, wgt) in ((y0, x0, hh * hw), (y0, x0 + 1, hh * lw),
                                  (y0 + 1, x0, lh * hw), (y0 + 1, x0 + 1, lh * lw)):
                m = ok & (yy >= 0) & (yy <= Hl - 1) & (xx >= 0) & (xx <= Wl - 1)
                idx = lsi[l] + np.clip(yy, 0, Hl - 1) * Wl + np.clip(xx, 0, Wl - 1)
                v = value[bi, idx, hi] * m[..., None]          # [B,Q,H,C]
                corners.append((idx, m, wgt, v))
            a = aw[:, :, :, l, p]
            val = sum(wgt[..., None] * v for (_, _, wgt, v) in corners)
            out += a[..., None] * val
            if grad_output is not None:
                v1, v2, v3, v4 = (c[3] for c in corners)
                top = go * a[..., None]
                for (idx, m, wgt, _) in corners:
                    contrib = (wgt * m)[..., None] * top
                    np.add.at(gv, (np.broadcast_to(bi, idx.shape), idx,
                                   np.broadcast_to(hi, idx.shape)), contrib)
                g_w = (hh[..., None] * (v2 - v1) + lh[..., None] * (v4 - v3))
                g_h = (hw[..., None] * (v3 - v1) + lw[..., None] * (v4 - v2))
                gloc[:, :, :, l, p, 0] = Wl * (g_w * top).sum(-1) * ok
                gloc[:, :, :, l, p, 1] = Hl * (g_h * top).sum(-1) * ok
                gaw[:, :, :, l, p] = (go * val).sum(-1) * ok
    out = out.reshape(B, Q, H * C)
    if grad_output is None:
        return out
    return out, gv, gloc, gaw
