"""ViDAR head ray operations on the GPU kernels of vidar_b200/csrc/ray_head.cu.

Host-side mirror of the ray part of
projects/mmdet3d_plugin/bevformer/dense_heads/vidar_head_base.py:

  get_grid_features(...)      == ViDARHeadBase._get_grid_features (:420-509): same arguments
                                 (plus ray_grid_num, an attribute of the head), same four
                                 returned tensors, same ray order and dropping rule;
  ce_regularization_loss(...) == the `use_ce_loss` branch of ViDARHeadBase.loss (:586-592),
                                 computed by the FUSED sampler+cross-entropy kernels: the
                                 [lvl, R, 513] logits are never materialised;
  decode_ray_depth(...)       == the per-(batch, frame) body of get_point_cloud_prediction
                                 (:706-738): arg-max waypoint -> distance.

Low-level autograd ops: `ray_sample` (logits/lengths, differentiable w.r.t. sigma) and `ray_ce`.
All tensors must be CUDA; there is no PyTorch fallback.
"""
import torch
from torch.autograd.function import once_differentiable

from . import _lib


def _prep(sigma, origin, points, frame):
    _lib.require_cuda(sigma=sigma, origin=origin, points=points, frame=frame)
    if sigma.dim() != 4 or origin.dim() != 2 or points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("expected sigma [F,Z,Y,X], origin [F,3], points [R,3]")
    if sigma.dtype != torch.float32 or origin.dtype != torch.float32 or points.dtype != torch.float32:
        raise RuntimeError("sigma / origin / points must be float32")
    F, Z, Y, X = sigma.shape
    if origin.shape[0] != F:
        raise RuntimeError("origin must hold one row per frame of sigma")
    if frame is not None and (frame.dtype != torch.int32 or frame.shape != (points.shape[0],)):
        raise RuntimeError("frame must be int32 [R]")
    return points.shape[0], F, Z, Y, X


def _call(fn, device, *args):
    with torch.cuda.device(device):
        _lib.check(fn(*args, _lib.stream_ptr(device)))


class _RaySample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, origin, points, frame, num_way, step, with_gt):
        sigma = sigma.contiguous()
        R, F, Z, Y, X = _prep(sigma, origin, points, frame)
        K = num_way + int(with_gt)
        logits = torch.empty((R, K), dtype=torch.float32, device=sigma.device)
        length = torch.empty((R, K), dtype=torch.float32, device=sigma.device)
        valid = torch.empty((R,), dtype=torch.float32, device=sigma.device)
        if R:
            _call(_lib.lib().vidar_ray_sample, sigma.device, _lib.ptr(sigma), _lib.ptr(origin),
                  _lib.ptr(points), _lib.ptr(frame), _lib.ptr(logits), _lib.ptr(length),
                  _lib.ptr(valid), R, F, Z, Y, X, int(num_way), float(step), int(with_gt))
        ctx.save_for_backward(origin, points, frame)
        ctx.meta = (tuple(sigma.shape), int(num_way), float(step), int(with_gt))
        ctx.mark_non_differentiable(length, valid)
        return logits, length, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_logits, _gl, _gv):
        origin, points, frame = ctx.saved_tensors
        shape, num_way, step, with_gt = ctx.meta
        F, Z, Y, X = shape
        grad_sigma = torch.zeros(shape, dtype=torch.float32, device=grad_logits.device)
        R = points.shape[0]
        if R:
            # -inf logits receive NaN/0 gradients from log-softmax; masked samples are skipped
            g = torch.nan_to_num(grad_logits.float(), nan=0.0, posinf=0.0, neginf=0.0).contiguous()
            _call(_lib.lib().vidar_ray_sample_backward, g.device, _lib.ptr(origin), _lib.ptr(points),
                  _lib.ptr(frame), _lib.ptr(g), _lib.ptr(grad_sigma), R, F, Z, Y, X, num_way, step, with_gt)
        return grad_sigma, None, None, None, None, None, None


def ray_sample(sigma, origin, points, frame, num_way, step, with_gt=True):
    """-> (logits [R,K], length [R,K], valid [R]); K = num_way + with_gt; differentiable in sigma."""
    return _RaySample.apply(sigma, origin.contiguous(), points.contiguous(), frame, num_way, step, with_gt)


class _RayCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, origin, points, frame, num_way, step):
        sigma = sigma.contiguous()
        R, F, Z, Y, X = _prep(sigma, origin, points, frame)
        ce = torch.zeros((R,), dtype=torch.float32, device=sigma.device)
        lse = torch.zeros((R,), dtype=torch.float32, device=sigma.device)
        valid = torch.zeros((R,), dtype=torch.float32, device=sigma.device)
        if R:
            _call(_lib.lib().vidar_ray_ce_forward, sigma.device, _lib.ptr(sigma), _lib.ptr(origin),
                  _lib.ptr(points), _lib.ptr(frame), _lib.ptr(ce), _lib.ptr(lse), _lib.ptr(valid),
                  R, F, Z, Y, X, int(num_way), float(step))
        ctx.save_for_backward(sigma, origin, points, frame, lse)
        ctx.meta = (int(num_way), float(step))
        ctx.mark_non_differentiable(valid)
        return ce, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_ce, _gv):
        sigma, origin, points, frame, lse = ctx.saved_tensors
        num_way, step = ctx.meta
        F, Z, Y, X = sigma.shape
        grad_sigma = torch.zeros_like(sigma)
        R = points.shape[0]
        if R:
            g = grad_ce.float().contiguous()
            _call(_lib.lib().vidar_ray_ce_backward, sigma.device, _lib.ptr(sigma), _lib.ptr(origin),
                  _lib.ptr(points), _lib.ptr(frame), _lib.ptr(lse), _lib.ptr(g), _lib.ptr(grad_sigma),
                  R, F, Z, Y, X, num_way, step)
        return grad_sigma, None, None, None, None, None


def ray_ce(sigma, origin, points, frame, num_way, step):
    """Fused sampler + cross-entropy(label 0) -> (ce [R], valid [R])."""
    return _RayCE.apply(sigma, origin.contiguous(), points.contiguous(), frame, num_way, step)


class _RayGumbel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, origin, points, frame, noise, num_way, step):
        sigma = sigma.contiguous()
        R, F, Z, Y, X = _prep(sigma, origin, points, frame)
        noise = noise.float().contiguous()
        if tuple(noise.shape) != (R, int(num_way)):
            raise RuntimeError(f"noise must be [rays, num_way] = {(R, int(num_way))}, got {tuple(noise.shape)}")
        out = torch.zeros((3, R), dtype=torch.float32, device=sigma.device)      # dist, lse, p_next
        if R:
            _call(_lib.lib().vidar_ray_gumbel_forward, sigma.device, _lib.ptr(sigma), _lib.ptr(origin),
                  _lib.ptr(points), _lib.ptr(frame), _lib.ptr(noise), _lib.ptr(out[0]), _lib.ptr(out[1]),
                  _lib.ptr(out[2]), R, F, Z, Y, X, int(num_way), float(step))
        ctx.save_for_backward(sigma, origin, points, frame, out)
        ctx.meta = (int(num_way), float(step))
        return out[0].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_dist):
        sigma, origin, points, frame, out = ctx.saved_tensors
        num_way, step = ctx.meta
        F, Z, Y, X = sigma.shape
        grad_sigma = torch.zeros_like(sigma)
        R = points.shape[0]
        if R:
            g = grad_dist.float().contiguous()
            _call(_lib.lib().vidar_ray_gumbel_backward, sigma.device, _lib.ptr(sigma), _lib.ptr(origin),
                  _lib.ptr(points), _lib.ptr(frame), _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]),
                  _lib.ptr(g), _lib.ptr(grad_sigma), R, F, Z, Y, X, num_way, step)
        return grad_sigma, None, None, None, None, None, None


def gumbel_distance(sigma, origin, points, frame, num_way, step, noise):
    """Fused sampler + `_custom_gumbel_softmax_distance` (vidar_head_base.py:754-773) over the
    num_way waypoints of every ray (no GT slot): sigma [F,Z,Y,X], origin [F,3], points [R,3],
    frame int32 [R], noise [R,num_way] Gumbel(0,1) -> dist [R], differentiable in sigma."""
    return _RayGumbel.apply(sigma, origin.contiguous(), points.contiguous(), frame, noise, num_way, step)


def ray_argmax(sigma, origin, points, frame, num_way, step):
    """-> (depth [R] voxel units, index [R] float) of the arg-max waypoint."""
    sigma = sigma.contiguous()
    origin, points = origin.contiguous(), points.contiguous()
    R, F, Z, Y, X = _prep(sigma, origin, points, frame)
    depth = torch.zeros((R,), dtype=torch.float32, device=sigma.device)
    index = torch.zeros((R,), dtype=torch.float32, device=sigma.device)
    if R:
        _call(_lib.lib().vidar_ray_argmax, sigma.device, _lib.ptr(sigma), _lib.ptr(origin),
              _lib.ptr(points), _lib.ptr(frame), _lib.ptr(depth), _lib.ptr(index),
              R, F, Z, Y, X, int(num_way), float(step))
    return depth, index


def _frame_sorted_rays(gt_grids, gt_tindex, num_frames):
    """Rays of one batch element grouped by frame (frame-major, original order inside a frame,
    like the reference's double loop) -> (points [R,3], frame int32 [R])."""
    t = gt_tindex
    keep = (t >= 0) & (t < num_frames) & (t == t.floor())     # `cur_tindex == frame_idx`
    idx = keep.nonzero().squeeze(-1)
    order = torch.argsort(t[idx], stable=True)
    idx = idx[order]
    return gt_grids[idx].contiguous(), t[idx].to(torch.int32).contiguous()


def get_grid_features(batched_origin_grids, batched_gt_grids, batched_gt_tindex, intermediate_sigma,
                      loss_weights, ray_grid_step, ray_grid_num, return_as_batch=False):
    """ViDARHeadBase._get_grid_features (:420-509).  intermediate_sigma: list (one per
    intermediate level) of [bs, F, Z, Y, X]; loss_weights indexable as [frame, level].
    Returns (r_mask_total [R,K] 0/-inf, r_feat_total [lvl,R,K], r_loss_weight_total [lvl,R],
    r_grid_length [R,K]) with invalid-GT rays dropped."""
    bs, F = intermediate_sigma[0].shape[:2]
    lw = torch.as_tensor(loss_weights, dtype=torch.float32, device=batched_gt_grids.device)
    masks, feats, weights, lengths = [], [], [], []
    for b in range(bs):
        pts, frame = _frame_sorted_rays(batched_gt_grids[b], batched_gt_tindex[b], F)
        origin = batched_origin_grids[b, :F].contiguous().float()
        lvl_feats, keep, length = [], None, None
        for sig in intermediate_sigma:
            logits, length, valid = ray_sample(sig[b], origin, pts, frame, ray_grid_num, ray_grid_step, True)
            keep = valid > 0
            lvl_feats.append(logits[keep])
        feats.append(torch.stack(lvl_feats, 0))
        length = length[keep]
        lengths.append(length)
        masks.append(torch.isinf(feats[-1][0]) & (feats[-1][0] < 0))
        weights.append(lw[frame[keep].long()].t().contiguous())           # [lvl, R]
    r_mask = torch.cat(masks, 0)
    r_mask_total = torch.zeros_like(r_mask, dtype=torch.float32).masked_fill(r_mask, float("-inf"))
    r_feat_total = torch.cat(feats, 1)
    r_loss_weight_total = torch.cat(weights, 1)
    r_grid_length = torch.cat(lengths, 0)
    if return_as_batch:
        n_lvl = len(intermediate_sigma)
        r_mask_total = r_mask_total.view(bs, -1, *r_mask_total.shape[1:])
        r_feat_total = r_feat_total.view(n_lvl, bs, -1, *r_feat_total.shape[2:])
        r_loss_weight_total = r_loss_weight_total.view(n_lvl, bs, -1, *r_loss_weight_total.shape[2:])
        r_grid_length = r_grid_length.view(bs, -1, *r_grid_length.shape[1:])
    return r_mask_total, r_feat_total, r_loss_weight_total, r_grid_length


def ce_regularization_loss(batched_origin_grids, batched_gt_grids, batched_gt_tindex, intermediate_sigma,
                           loss_weights, ray_grid_step, ray_grid_num):
    """`regularization.loss` of ViDARHeadBase.loss (:586-592) without materialising the logits:
    sum(ce * w) / clamp(sum(w), min=1) over intermediate levels and valid rays."""
    bs, F = intermediate_sigma[0].shape[:2]
    lw = torch.as_tensor(loss_weights, dtype=torch.float32, device=batched_gt_grids.device)
    num = batched_gt_grids.new_zeros(())
    den = batched_gt_grids.new_zeros(())
    for b in range(bs):
        pts, frame = _frame_sorted_rays(batched_gt_grids[b], batched_gt_tindex[b], F)
        origin = batched_origin_grids[b, :F].contiguous().float()
        for lvl, sig in enumerate(intermediate_sigma):
            ce, valid = ray_ce(sig[b], origin, pts, frame, ray_grid_num, ray_grid_step)
            w = lw[frame.long(), lvl] * valid
            num = num + (ce * w).sum()
            den = den + w.sum()
    return num / torch.clamp(den, min=1)


def decode_ray_depth(sigma, origin_grids, gt_grids, gt_tindex, ray_grid_step, ray_grid_num):
    """Per batch element: sigma [F,Z,Y,X], origin_grids [F,3], gt_grids [M,3], gt_tindex [M] ->
    (pred_dist [M], gt_dist [M]) in voxel units, zeros for rays of no frame
    (get_point_cloud_prediction :700-734)."""
    F = sigma.shape[0]
    t = gt_tindex
    keep = (t >= 0) & (t < F) & (t == t.floor())
    idx = keep.nonzero().squeeze(-1)
    frame = t[idx].to(torch.int32).contiguous()
    pts = gt_grids[idx].contiguous()
    depth, _ = ray_argmax(sigma, origin_grids[:F].contiguous().float(), pts, frame, ray_grid_num, ray_grid_step)
    pred = gt_grids.new_zeros(gt_grids.shape[0])
    gt = gt_grids.new_zeros(gt_grids.shape[0])
    pred[idx] = depth
    gt[idx] = torch.sqrt(((pts - origin_grids[frame.long()]) ** 2).sum(-1))
    return pred, gt
