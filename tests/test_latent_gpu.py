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
GOLD_FUSED = os.path.join(os.path.dirname(__file__), "golden", "latent_fused.npz")


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


def _run_module(m, c, cuda):
    e = c["embed"].to(cuda).requires_grad_(True)
    m.zero_grad(set_to_none=True)
    out = m(e)
    out.backward(c["grad"].to(cuda))
    return out.detach(), e.grad, {n: p.grad.clone() for n, p in m.named_parameters()}


@pytest.mark.parametrize("tag,cfg,seed", [("fused", lc.CFG_FUSED, 23), ("fused_exp", lc.CFG_FUSED_EXP, 24)])
def test_fused_module_matches_reference_class_and_unfused_path(cuda, tag, cfg, seed):
    """csrc/latent_proj.cu: the module as six kernels (projections fused around the ray-marching
    core) against the REFERENCE class's outputs and every gradient, and against this package's own
    unfused path (cuBLAS Linears around the same core)."""
    from vidar_b200 import _lib
    g = np.load(GOLD_FUSED)
    m = build_attention(cfg)
    m.load_state_dict(lc.seeded_state(m, seed))
    m.to(cuda)
    c = lc.case(seed=5, bev=lc.BEV_FUSED, embed_dims=cfg["embed_dims"])
    n0 = _lib.launch_count()
    out, ge, gp = _run_module(m, c, cuda)
    assert _lib.launch_count() - n0 == 8          # 4 projection + 4 ray-marching kernels, nothing else of ours
    _close(out, g[f"{tag}_out"], "output")
    _close(ge, g[f"{tag}_gembed"], "grad embed", rtol=2e-4, frac=2e-4)
    for n, v in gp.items():
        _close(v, g[f"{tag}_g_{n}"], f"grad {n}", rtol=2e-4, frac=2e-4)
    m.fuse_projections = False
    out2, ge2, gp2 = _run_module(m, c, cuda)
    _close(out, out2, "fused vs unfused output", rtol=2e-5, frac=2e-5)
    _close(ge, ge2, "fused vs unfused grad embed", rtol=1e-4, frac=1e-4)
    for n in gp:
        _close(gp[n], gp2[n], f"fused vs unfused grad {n}", rtol=1e-4, frac=1e-4)


@pytest.mark.parametrize("rows,E,D", [(1, 256, 16), (31, 256, 16), (1000, 256, 16), (4099, 128, 8), (777, 256, 4),
                                      (0, 256, 16)])
def test_projection_kernels_vs_torch(cuda, rows, E, D):
    """The four projection kernels one by one against the torch fp32 formulas (cuBLAS, no TF32),
    ragged row counts (tile tails), both widths, three height groupings."""
    import torch.nn.functional as F
    from vidar_b200 import _lib
    A = 16
    g = torch.Generator().manual_seed(rows + E + D)
    rnd = lambda *s: torch.randn(*s, generator=g).to(cuda)
    x, w_occ, b_occ, w_feat, b_feat = rnd(rows, E), rnd(D, E) / 16, rnd(D), rnd(A, E) / 16, rnd(A)
    w_b, b_b, pooled, prob = rnd(E, A) / 4, rnd(E), rnd(rows, A), torch.rand(rows, D, generator=g).to(cuda)
    g_occ, g_feat, g_out = rnd(rows, D), rnd(rows, A), rnd(rows, E)
    L, st = _lib.lib(), _lib.stream_ptr(cuda)
    P = _lib.ptr
    occ, feat = torch.empty(rows, D, device=cuda), torch.empty(rows, A, device=cuda)
    _lib.check(L.vidar_latent_proj_in_forward(P(x), P(w_occ), P(b_occ), P(w_feat), P(b_feat), P(occ), P(feat), rows, E, D, A, st))
    out = torch.empty(rows, E, device=cuda)
    _lib.check(L.vidar_latent_proj_out_forward(P(pooled), P(prob), P(w_b), P(b_b), P(out), rows, E, D, A, st))
    gx = torch.empty(rows, E, device=cuda)
    gwo, gbo, gwf, gbf = (torch.zeros_like(t) for t in (w_occ, b_occ, w_feat, b_feat))
    _lib.check(L.vidar_latent_proj_in_backward(P(x), P(w_occ), P(w_feat), P(g_occ), P(g_feat), P(gx), P(gwo), P(gbo), P(gwf),
                                               P(gbf), rows, E, D, A, st))
    gpo, gpr = torch.empty(rows, A, device=cuda), torch.empty(rows, D, device=cuda)
    gwb, gbb = torch.zeros_like(w_b), torch.zeros_like(b_b)
    _lib.check(L.vidar_latent_proj_out_backward(P(g_out), P(pooled), P(prob), P(w_b), P(b_b), P(gpo), P(gpr), P(gwb), P(gbb),
                                                rows, E, D, A, st))
    if rows == 0:
        assert float(gwo.abs().sum() + gwb.abs().sum()) == 0
        return
    # torch reference in fp64 (the kernels and cuBLAS both round differently from it by ~1e-6)
    dd = lambda t: t.double().requires_grad_(True)
    X, WO, BO, WF, BF, WB, BB, PO, PR = map(dd, (x, w_occ, b_occ, w_feat, b_feat, w_b, b_b, pooled, prob))
    r_occ, r_feat = F.linear(X, WO, BO), F.linear(X, WF, BF)
    r_out = (F.linear(PO, WB, BB).view(rows, D, E // D) * PR.view(rows, D, 1)).view(rows, E)
    (r_occ * g_occ.double()).sum().backward(retain_graph=True)
    (r_feat * g_feat.double()).sum().backward()
    (r_out * g_out.double()).sum().backward()
    for a, b, w in ((occ, r_occ, "occ"), (feat, r_feat, "feat"), (out, r_out, "out"), (gx, X.grad, "grad embed"),
                    (gwo, WO.grad, "grad w_occ"), (gbo, BO.grad, "grad b_occ"), (gwf, WF.grad, "grad w_feat"),
                    (gbf, BF.grad, "grad b_feat"), (gpo, PO.grad, "grad pooled"), (gpr, PR.grad, "grad prob"),
                    (gwb, WB.grad, "grad w_b"), (gbb, BB.grad, "grad b_b")):
        _close(a, b, w, rtol=2e-5, frac=1e-5)


def test_projection_kernels_reject_unsupported_shapes(cuda):
    from vidar_b200 import _lib
    t = torch.zeros(64, device=cuda)
    L, st, P = _lib.lib(), _lib.stream_ptr(cuda), _lib.ptr
    with pytest.raises(RuntimeError, match="embed_dims must be 128 or 256"):
        _lib.check(L.vidar_latent_proj_in_forward(P(t), P(t), P(t), P(t), P(t), P(t), P(t), 1, 64, 16, 16, st))
    with pytest.raises(RuntimeError, match="rank"):
        _lib.check(L.vidar_latent_proj_out_backward(P(t), P(t), P(t), P(t), P(t), P(t), P(t), P(t), P(t), 1, 256, 8, 8, st))
