"""Work partitioning of the hot path over ranks (one process per GPU).

SURVEY.md 8(e): the MSDA stage shards the (camera, query) rows of SpatialCrossAttention --
cameras are independent until the scatter-add into the BEV slots
(spatial_cross_attention.py:164-171) -- and exchanges the per-row outputs ONCE (all-gather of
the BEV rows); the ray stages shard rays and all-reduce the small grad_sigma volume.  The
reference has no counterpart (replica DDP only, apis/mmdet_train.py:72-81).
Pure host logic: usable with any torch.distributed backend (tests run it over gloo on CPU).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) share of n items."""
    return rank * n // world, (rank + 1) * n // world


def shard_rows(rank, world, cams, Q):
    """Contiguous share of the cams*Q (camera, query) rows -> [(cam, q0, q1), ...]."""
    lo, hi = shard_range(cams * Q, rank, world)
    segs = []
    for c in range(cams):
        a, b = max(lo, c * Q), min(hi, (c + 1) * Q)
        if a < b:
            segs.append((c, a - c * Q, b - c * Q))
    return segs


def cameras_of(rank, world, cams, Q):
    return sorted({c for c, _, _ in shard_rows(rank, world, cams, Q)})


def camera_groups(world, cams, Q, rank):
    """Process groups for cameras whose rows are split over several ranks (their partial
    grad_value must be summed).  EVERY rank must call this (new_group is collective).
    -> {cam: group} for the cameras this rank shares."""
    groups = {}
    for c in range(cams):
        members = [r for r in range(world) if c in cameras_of(r, world, cams, Q)]
        if len(members) > 1:
            g = dist.new_group(members)
            if rank in members:
                groups[c] = g
    return groups


def gather_rows(local_rows, world, total_rows, group=None):
    """all-gather of per-rank row blocks [n_r, C] (balanced contiguous shards) -> [total_rows, C]."""
    if world == 1:
        return local_rows
    if total_rows % world == 0:
        full = local_rows.new_empty(total_rows, local_rows.shape[1])
        dist.all_gather_into_tensor(full.view(-1), local_rows.contiguous().view(-1), group=group)
        return full
    max_rows = -(-total_rows // world)
    pad = local_rows.new_zeros(max_rows, local_rows.shape[1])
    pad[: local_rows.shape[0]] = local_rows
    flat = local_rows.new_empty(world * max_rows, local_rows.shape[1])
    dist.all_gather_into_tensor(flat.view(-1), pad.view(-1), group=group)
    buf = flat.view(world, max_rows, local_rows.shape[1])
    parts = []
    for r in range(world):
        lo, hi = shard_range(total_rows, r, world)
        parts.append(buf[r, : hi - lo])
    return torch.cat(parts, 0)
