"""LatentRendering: CPU oracle + module host logic against goldens from the REFERENCE class
(tools/make_golden_latent.py).  The CUDA core is replaced by the oracle core here."""
import os

import numpy as np
import pytest
import torch

from oracle import latent_render_ref as ref
from tests import latent_cases as lc
from vidar_b200.modules import latent_rendering as lr
from vidar_b200.registry import build_attention

GOLD = os.path.join(os.path.dirname(__file__), "golden", "latent_rendering.npz")
GOLD_FUSED = os.path.join(os.path.dirname(__file__), "golden", "latent_fused.npz")


@pytest.fixture()
def oracle_core(monkeypatch):
    def core(occ, feat, grid_num, grid_step, eps, act, group=None):
        return ref.latent_core(occ, feat, grid_num, grid_step, eps, "sigmoid" if act == 1 else "exp")
    monkeypatch.setattr(lr, "latent_render_core", core)


@pytest.mark.parametrize("tag,cfg,seed", [("sig", lc.CFG, 20), ("exp", lc.CFG_EXP, 21), ("d1", lc.CFG_D1, 22)])
def test_module_with_oracle_core_matches_reference_class(oracle_core, tag, cfg, seed):
    g = np.load(GOLD)
    m = build_attention(cfg)
    assert sorted(m.state_dict().keys()) == list(g[f"{tag}_params"])
    m.load_state_dict(lc.seeded_state(m, seed))
    c = lc.case()
    e = c["embed"].clone().requires_grad_(True)
    out = m(e)
    out.backward(c["grad"])
    np.testing.assert_allclose(out.detach().numpy(), g[f"{tag}_out"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(e.grad.numpy(), g[f"{tag}_gembed"], rtol=1e-3, atol=1e-6)
    for n, p in m.named_parameters():
        key = f"{tag}_g_{n}"
        if key in g:
            np.testing.assert_allclose(p.grad.numpy(), g[key], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("tag,cfg,seed", [("fused", lc.CFG_FUSED, 23), ("fused_exp", lc.CFG_FUSED_EXP, 24)])
def test_oracle_reproduces_goldens_of_the_fused_shapes(oracle_core, tag, cfg, seed):
    """The goldens the fused-projection GPU path is held to (tests/test_latent_gpu.py), checked here
    through the unfused host path + oracle core: outputs, grad embed and EVERY parameter gradient."""
    g = np.load(GOLD_FUSED)
    m = build_attention(cfg)
    assert sorted(m.state_dict().keys()) == list(g[f"{tag}_params"])
    assert lr.fused_projection_supported(m.embed_dims, m.pred_height, m.lora_a.out_features)
    m.load_state_dict(lc.seeded_state(m, seed))
    c = lc.case(seed=5, bev=lc.BEV_FUSED, embed_dims=cfg["embed_dims"])
    e = c["embed"].clone().requires_grad_(True)
    out = m(e)
    out.backward(c["grad"])
    np.testing.assert_allclose(out.detach().numpy(), g[f"{tag}_out"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(e.grad.numpy(), g[f"{tag}_gembed"], rtol=1e-3, atol=1e-6)
    for n, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), g[f"{tag}_g_{n}"], rtol=1e-3, atol=2e-6)


def test_fused_projection_shape_gate():
    ok = lr.fused_projection_supported
    assert ok(256, 16, 16) and ok(128, 8, 16) and ok(256, 4, 16)
    assert not ok(64, 16, 16)          # embed_dims not 128/256
    assert not ok(256, 16, 32)         # rank != 16
    assert not ok(256, 1, 16)          # 1 + 16 is not a multiple of 4 (class default pred_height)
    assert not ok(256, 32, 16)         # rank not a multiple of pred_height


def test_no_cpu_fallback():
    m = build_attention(lc.CFG)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        m(lc.case()["embed"])
