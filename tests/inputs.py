"""Seeded synthetic inputs shared by tests, smoke() and bench.py."""
import numpy as np
import torch

from vidar_b200.synthetic import dvr_inputs_lidar, level_tensors  # noqa: F401  (re-exported)

SCA_LEVELS = ((116, 200), (58, 100), (29, 50), (15, 25))   # 928x1600 input, strides 8..64


def msda_inputs(B, Q, H, C, levels, P, seed=0, mode="local", device="cpu", dtype=torch.float32):
    """mode 'local': loc = per-query reference point + N(0, 4px) offsets (like real init,
    spatial_cross_attention.py:255-266); 'stress': U(-0.1, 1.1) incl. out-of-range."""
    g = torch.Generator().manual_seed(seed)
    L = len(levels)
    K = sum(h * w for h, w in levels)
    value = torch.randn(B, K, H, C, generator=g, dtype=dtype)
    if mode == "local":
        ref = torch.rand(B, Q, 1, 1, 1, 2, generator=g, dtype=dtype)
        wh = torch.tensor([[w, h] for h, w in levels], dtype=dtype).view(1, 1, 1, L, 1, 2)
        loc = ref + 4.0 * torch.randn(B, Q, H, L, P, 2, generator=g, dtype=dtype) / wh
    else:
        loc = torch.rand(B, Q, H, L, P, 2, generator=g, dtype=dtype) * 1.2 - 0.1
    attn = torch.softmax(torch.randn(B, Q, H, L * P, generator=g, dtype=dtype), -1).view(B, Q, H, L, P)
    grad_out = torch.randn(B, Q, H * C, generator=g, dtype=dtype)
    shapes, lsi = level_tensors(levels)
    to = lambda t: t.to(device)
    return dict(value=to(value), shapes=to(shapes), lsi=to(lsi), loc=to(loc.contiguous()),
                attn=to(attn.contiguous()), grad_out=to(grad_out))


def dvr_inputs_cfg1(seed=0, integer_origin=True, M=1000, pad=16):
    """BASELINE.json configs[0]: 50x50x8 volume, 1k rays, fp32."""
    rng = np.random.default_rng(seed)
    sigma = rng.uniform(0, 1, (1, 1, 8, 50, 50)).astype(np.float32)
    o = (25.0, 25.0, 4.0) if integer_origin else (24.37, 25.61, 3.52)
    origin = np.array(o, np.float32).reshape(1, 1, 3)
    lo = np.array([-5, -5, -1], np.float32)
    hi = np.array([55, 55, 9], np.float32)
    points = (rng.uniform(0, 1, (1, M, 3)) * (hi - lo) + lo).astype(np.float32)
    tindex = np.zeros((1, M), np.float32)
    if pad:
        points[0, -pad:] = np.nan
        tindex[0, -pad:] = -1
    return sigma, origin, points, tindex


def dvr_inputs_outside(seed=11, M=800, zero_length=True):
    """Origins OUTSIDE the volume (the warp-per-ray kernels hand these to their serial path): rays
    that enter, rays that never enter, rays that end before entering and -- with `zero_length` --
    degenerate zero-length rays.  Frame 0's z origin has a .5 fraction, which puts the rounded-path
    variants' z crossings exactly on round() ties (the FMA in the path advance decides them).
    The reference itself never terminates on a zero-length ray that starts outside (dvr.cu:188-227:
    `last_d > NaN` is never true), so the golden fixture is generated with zero_length=False."""
    rng = np.random.default_rng(seed)
    sigma = rng.uniform(0, 1, (1, 2, 6, 30, 28)).astype(np.float32)
    origin = np.array([[[-7.3, 12.2, 2.5], [14.1, 40.6, 9.7]]], np.float32)       # frame 0: x<0 ; frame 1: y,z beyond
    points = (rng.uniform(0, 1, (1, M, 3)) * np.array([40, 44, 10]) - np.array([6, 7, 2])).astype(np.float32)
    tindex = rng.integers(0, 2, (1, M)).astype(np.float32)
    points[0, :5] = origin[0, tindex[0, :5].astype(int)]        # zero-length rays (NaN direction)
    points[0, 5:25] = origin[0, tindex[0, 5:25].astype(int)] + rng.normal(0, 0.5, (20, 3)).astype(np.float32)  # end before entering
    tindex[0, -9:] = -1
    if not zero_length:
        tindex[0, :5] = -1
    return sigma, origin, points, tindex


def dvr_inputs_ties(seed=123, M=6000):
    """Tie-prone rays: integer, HALF-INTEGER (every crossing is a round() tie of the rounded-path variants)
    and generic origins; a share of the end points on the integer / half-integer lattice (crossing times of
    two axes coincide).  Branch decisions on such rays hang on the last bits of the traversal arithmetic."""
    rng = np.random.default_rng(seed)
    Z, Y, X = 8, 50, 50
    sigma = rng.uniform(0, 1, (1, 3, Z, Y, X)).astype(np.float32)
    origin = np.array([[[25.0, 25.0, 4.0], [24.5, 25.5, 3.5], [24.37, 25.61, 3.52]]], np.float32)
    points = (rng.uniform(0, 1, (1, M, 3)) * np.array([60, 60, 10]) - np.array([5, 5, 1])).astype(np.float32)
    points[0, ::7] = np.round(points[0, ::7])
    points[0, ::11] = np.round(points[0, ::11] * 2) / 2
    tindex = rng.integers(0, 3, (1, M)).astype(np.float32)
    return sigma, origin, points, tindex
