"""LatentRendering on the GPU (fused CUDA core) against goldens from the REFERENCE class and
against the CPU oracle at the BASELINE size (200x200x16, 256 waypoints)."""
import os

import numpy as np
import pytest
import torch

from oracle import latent_render_ref as ref
from tests import latent_cases as lc
from vidar_b200.modules.latent_rendering import latent_render_core
import vidar_b200.modules  # noqa: F401  (registers the classes)
from vidar_b200.registry import build_attention

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "latent_rendering.npz")


def _close(a, b, what, rtol=1e-4, frac=1e-4):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    err = np.abs(a - b)
    assert err.max() <= rtol * np.abs(b).max() + 1e-30, f"{what}: max err {err.max():.3e} vs max|ref| {np.abs(b).max():.3e}"
    assert np.linalg.norm(a - b) <= frac * np.linalg.norm(b) + 1e-30, f"{what}: L2 err {np.linalg.norm(a-b):.3e} vs {np.linalg.norm(b):.3e}"


@pytest.mark.parametrize("tag,cfg,seed", [("sig", lc.CFG, 20), ("exp", lc.CFG_EXP, 21), ("d1", lc.CFG_D1, 22)])
def test_module_on_gpu_matches_reference_class(cuda, tag, cfg, seed):
    g = np.load(GOLD)
    m = build_attention(cfg)
    m.load_state_dict(lc.seeded_state(m, seed))
    m.to(cuda)
    c = lc.case()
    e = c["embed"].to(cuda).requires_grad_(True)
    out = m(e)
    out.backward(c["grad"].to(cuda))
    _close(out, g[f"{tag}_out"], "output")
    _close(e.grad, g[f"{tag}_gembed"], "grad embed", rtol=2e-4, frac=2e-4)
    for n, p in m.named_parameters():
        key = f"{tag}_g_{n}"
        if key in g:
            _close(p.grad, g[key], f"grad {n}", rtol=2e-4, frac=2e-4)


def test_core_full_size_vs_oracle(cuda):
    """200x200 BEV, 16 heights, 256 waypoints (BASELINE configs[2]b).  The oracle materialises
    [1,16,40000,257] tensors: run it on CPU once, compare everything."""
    g = torch.Generator().manual_seed(0)
    occ = torch.randn(1, 200, 200, 16, generator=g)
    feat = torch.randn(1, 200, 200, 16, generator=g)
    gp = torch.randn(1, 200, 200, 16, generator=g)
    gq = torch.randn(1, 40000, 16, generator=g)
    o_c, f_c = occ.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    rp, rq = ref.latent_core(o_c, f_c, 256, 0.5, 1e-3, "sigmoid")
    (rp * gp).sum().backward(retain_graph=True)
    (rq * gq).sum().backward()
    o_g, f_g = occ.to(cuda).requires_grad_(True), feat.to(cuda).requires_grad_(True)
    p, q = latent_render_core(o_g, f_g, 256, 0.5, 1e-3, 1)
    ((p * gp.to(cuda)).sum() + (q * gq.to(cuda)).sum()).backward()
    _close(p, rp, "prob")
    _close(q, rq, "pooled")
    _close(o_g.grad, o_c.grad, "grad occ", rtol=2e-4, frac=2e-4)
    _close(f_g.grad, f_c.grad, "grad feat", rtol=2e-4, frac=2e-4)
