"""Seeded inputs / weights for the module-level goldens (shared by tools/make_golden_modules.py
and the tests, so only outputs need to be stored)."""
import torch

E, HEADS = 256, 8
CAM_LEVELS = ((12, 20), (6, 10), (3, 5), (2, 3))
BEV = (10, 10)


def seeded_state(module, seed):
    """Deterministic weights for every parameter, keyed by name order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, p in sorted(module.state_dict().items()):
        scale = 0.5 if name.endswith("bias") else 1.0 / (p.shape[-1] ** 0.5)
        if "sampling_offsets.bias" in name:
            scale = 3.0        # a few pixels, like the trained models
        sd[name] = torch.randn(p.shape, generator=g) * scale
    return sd


def levels(shapes):
    s = torch.tensor(shapes, dtype=torch.int64)
    hw = s[:, 0] * s[:, 1]
    return s, torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])


def sca_case(seed=0, bs=1, cams=6, D=4):
    g = torch.Generator().manual_seed(seed)
    Q = BEV[0] * BEV[1]
    K = sum(h * w for h, w in CAM_LEVELS)
    shapes, lsi = levels(CAM_LEVELS)
    ref_cam = torch.rand(cams, bs, Q, D, 2, generator=g) * 1.3 - 0.15
    bev_mask = torch.rand(cams, bs, Q, D, generator=g) < 0.3
    bev_mask[3] = False                      # a camera that sees nothing
    bev_mask[:, :, 7] = False                # a pillar no camera sees
    return dict(query=torch.randn(bs, Q, E, generator=g), query_pos=0.1 * torch.randn(bs, Q, E, generator=g),
                key=torch.randn(cams, K, bs, E, generator=g), reference_points_cam=ref_cam, bev_mask=bev_mask,
                spatial_shapes=shapes, level_start_index=lsi, grad=torch.randn(bs, Q, E, generator=g))


def tsa_case(seed=1, bs=1):
    g = torch.Generator().manual_seed(seed)
    Q = BEV[0] * BEV[1]
    shapes, lsi = levels((BEV,))
    return dict(query=torch.randn(bs, Q, E, generator=g), query_pos=0.1 * torch.randn(bs, Q, E, generator=g),
                value=torch.randn(bs * 2, Q, E, generator=g),
                reference_points=torch.rand(bs * 2, Q, 1, 2, generator=g),
                spatial_shapes=shapes, level_start_index=lsi, grad=torch.randn(bs, Q, E, generator=g))


def pred_case(seed=2, bs=1, frames=2):
    g = torch.Generator().manual_seed(seed)
    Q = BEV[0] * BEV[1]
    shapes, lsi = levels((BEV,) * frames)
    return dict(query=torch.randn(bs, Q, E, generator=g), query_pos=0.1 * torch.randn(bs, Q, E, generator=g),
                value=torch.randn(bs, Q * frames, E, generator=g),
                reference_points=torch.rand(bs, Q, frames, 2, generator=g),
                spatial_shapes=shapes, level_start_index=lsi, grad=torch.randn(bs, Q, E, generator=g))


def custom_case(seed=3, bs=2, boxes=False, nq=30):
    """Detection-decoder layout: sequence-first query/value, one BEV level; `boxes` uses the
    4-component reference boxes (cx, cy, w, h) branch (decoder.py:317-321)."""
    g = torch.Generator().manual_seed(seed)
    K = BEV[0] * BEV[1]
    shapes, lsi = levels((BEV,))
    ref = torch.rand(bs, nq, 1, 2, generator=g)
    if boxes:
        ref = torch.cat([ref, 0.1 + 0.3 * torch.rand(bs, nq, 1, 2, generator=g)], -1)
    return dict(query=torch.randn(nq, bs, E, generator=g), query_pos=0.1 * torch.randn(nq, bs, E, generator=g),
                value=torch.randn(K, bs, E, generator=g), reference_points=ref,
                spatial_shapes=shapes, level_start_index=lsi, grad=torch.randn(nq, bs, E, generator=g))


SCA_CFG = dict(type="SpatialCrossAttention", pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], dropout=0.0,
               deformable_attention=dict(type="MSDeformableAttention3D", embed_dims=E, num_points=8, num_levels=4),
               embed_dims=E)
TSA_CFG = dict(type="TemporalSelfAttention", embed_dims=E, num_levels=1, dropout=0.0)
PRED_CFG = dict(type="PredictionMSDeformableAttention", embed_dims=E, num_levels=2, dropout=0.0)
CUSTOM_CFG = dict(type="CustomMSDeformableAttention", embed_dims=E, num_levels=1, dropout=0.0)


def run_module(m, kind, case, device="cpu"):
    """forward + backward of one module; returns (out, grad wrt query, grad wrt key/value)."""
    c = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in case.items()}
    q = c["query"].clone().requires_grad_(True)
    if kind == "sca":
        kv = c["key"].clone().requires_grad_(True)
        out = m(q, kv, kv, query_pos=c["query_pos"], reference_points_cam=c["reference_points_cam"],
                bev_mask=c["bev_mask"], spatial_shapes=c["spatial_shapes"], level_start_index=c["level_start_index"])
    else:
        kv = c["value"].clone().requires_grad_(True)
        out = m(q, None, kv, query_pos=c["query_pos"], reference_points=c["reference_points"],
                spatial_shapes=c["spatial_shapes"], level_start_index=c["level_start_index"])
    out.backward(c["grad"])
    return out.detach(), q.grad, kv.grad
