"""Seeded synthetic workloads of the hot path: the BASELINE.json configurations as
tensors.  Used by bench.py, __graft_entry__.smoke(), tools/ and the tests -- there is no dataset
and no checkpoint in this tier, so the workload generators are part of the package.

  sca_like_inputs    configs[1]: MSDA operands at the SpatialCrossAttention shape
  dvr_inputs_lidar   configs[2]: sigma volume + LiDAR-like rays
  camera_rig         cfg4: a nuScenes-like 6-camera rig (lidar2img matrices) for point_sampling
"""
import math

import numpy as np
import torch

SCA_LEVELS = ((116, 200), (58, 100), (29, 50), (15, 25))   # 928x1600 input, strides 8..64
IMG_HW = (928, 1600)


def level_tensors(levels, device="cpu"):
    shapes = torch.tensor(levels, dtype=torch.int64, device=device)
    hw = shapes[:, 0] * shapes[:, 1]
    lsi = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
    return shapes, lsi


def sca_like_inputs(device, cams=6, Q=40000, seed=0, levels=SCA_LEVELS, heads=8, head_dim=32, points=8,
                    rows=None):
    """Every camera sees a fan of Q pillars of its own frustum: perspective projection of a
    sqrt(Q) x sqrt(Q) polar BEV patch with 4 Z-anchors (-4,-2,0,2 m, camera at 1.5 m), f = 1266 px
    on a 1600x928 image, plus N(0, 4 px) learned-offset noise per (head, level, point)
    (SURVEY.md 8d cfg2).  Bottom anchors of near pillars fall outside the image, as in the real rig.
    `rows`: keep only the first `rows` queries of every camera (the "rebatched" mode, 10240)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L, P, H = len(levels), points, heads
    K = sum(h * w for h, w in levels)
    n = int(math.isqrt(Q))
    assert n * n == Q
    iy, ix = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
    depth = 2.0 + 49.0 * (iy.reshape(-1).float() + 0.5) / n               # 2 .. 51 m
    ang = ((ix.reshape(-1).float() + 0.5) / n - 0.5) * math.radians(60.0)
    u = 0.5 + torch.tan(ang) * 1266.0 / 1600.0
    zs = torch.tensor([-4.0, -2.0, 0.0, 2.0])
    v = (491.0 + 1266.0 * (1.5 - zs)[None, :] / depth[:, None]) / 928.0    # [Q, 4]
    ref = torch.stack([u[:, None].expand(-1, 4), v], -1)                   # [Q, 4(z), 2]
    ref = ref[None].expand(cams, -1, -1, -1).clone()
    ref += 0.01 * torch.randn(cams, 1, 1, 2, generator=g)                  # per-camera jitter
    if rows is not None:
        # rebatched mode: a camera keeps an evenly strided subset of its pillars
        keep = torch.linspace(0, Q - 1, rows).round().long()
        ref, Q = ref[:, keep].contiguous(), rows
    ref = ref.to(device)
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32, device=device)
    dg = torch.Generator(device=device).manual_seed(seed + 1)
    off = 4.0 * torch.randn(cams, Q, H, L, P, 2, device=device, generator=dg) / wh.view(1, 1, 1, L, 1, 2)
    # point p = j*4 + z uses Z-anchor z (spatial_cross_attention.py:356-371)
    loc = off.view(cams, Q, H, L, P // 4, 4, 2) + ref.view(cams, Q, 1, 1, 1, 4, 2)
    loc = loc.view(cams, Q, H, L, P, 2).contiguous()
    attn = torch.softmax(torch.randn(cams, Q, H, L * P, device=device, generator=dg), -1)
    attn = attn.view(cams, Q, H, L, P).contiguous()
    value = torch.randn(cams, K, H, head_dim, device=device, generator=dg)
    grad_out = torch.randn(cams, Q, H * head_dim, device=device, generator=dg)
    shapes, lsi = level_tensors(levels, device)
    return dict(value=value, shapes=shapes, lsi=lsi, loc=loc, attn=attn, grad_out=grad_out)


def dvr_inputs_lidar(M=30000, T=3, grid=(16, 200, 200), seed=0, N=1, pad=0):
    """BASELINE.json configs[2]: sigma [N,T,16,200,200] = softplus(N(0,1)), origin ~ grid centre,
    32-beam LiDAR-like endpoints, range U(2,70) m at 0.512 m/voxel."""
    rng = np.random.default_rng(seed)
    Z, Y, X = grid
    sigma = np.log1p(np.exp(rng.standard_normal((N, T, Z, Y, X)))).astype(np.float32)
    origin = (np.array([X / 2, Y / 2, Z * 5.0 / 8.0]) + rng.normal(0, 0.5, (N, T, 3))).astype(np.float32)
    beams = np.deg2rad(np.linspace(-30, 10, 32))
    elev = beams[rng.integers(0, 32, (N, M))]
    azim = rng.uniform(-np.pi, np.pi, (N, M))
    rng_m = rng.uniform(2, 70, (N, M)) / 0.512
    tindex = (np.arange(M) * T // M).astype(np.float32)[None].repeat(N, 0)
    d = np.stack([np.cos(elev) * np.cos(azim), np.cos(elev) * np.sin(azim),
                  np.sin(elev) * (Z / 8.0) / (X / 102.4)], -1)  # z voxels are 0.5 m, x/y 0.512 m
    o_per_ray = np.take_along_axis(origin, tindex.astype(np.int64)[..., None].repeat(3, -1), 1)
    points = (o_per_ray + d * rng_m[..., None]).astype(np.float32)
    if pad:
        points[:, -pad:] = np.nan
        tindex[:, -pad:] = -1
    return sigma, origin, points, tindex


def camera_rig(cams=6, img_hw=(900, 1600), f=1266.0, pp=(816.0, 491.0), cam_height=1.5, seed=0):
    """lidar2img [cams,4,4] (float32) of a nuScenes-like rig: yaw
    {0, +-55, +-110, 180} deg, focal 1266 px, principal point (816, 491), cameras 1.5 m above the
    LiDAR origin plane and ~1 m out from the centre, so each BEV pillar projects into 1-2 cameras."""
    yaws = np.deg2rad([0.0, 55.0, -55.0, 110.0, -110.0, 180.0])[:cams]
    rng = np.random.default_rng(seed)
    K = np.array([[f, 0, pp[0], 0], [0, f, pp[1], 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    # camera axes in the LiDAR frame: z forward, x right, y down
    mats = []
    for yaw in yaws:
        yaw = yaw + rng.normal(0, 0.01)
        fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
        right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd], 0)                 # rows: camera axes
        pos = fwd * 1.0 + np.array([0, 0, cam_height])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = -R @ pos
        mats.append(K @ E)
    return torch.from_numpy(np.stack(mats, 0).astype(np.float32))
