"""The attention modules end to end on the GPU (CUDA MSDA op inside) against goldens produced
by the REFERENCE classes on CPU (tools/make_golden_modules.py).  1e-4 relative fp32."""
import os

import numpy as np
import pytest
import torch

from tests import module_cases as mc
import vidar_b200.modules  # noqa: F401  (registers the classes)
from vidar_b200.registry import build_attention

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "modules.npz")
GOLD_CUSTOM = os.path.join(os.path.dirname(__file__), "golden", "modules_custom.npz")


@pytest.mark.parametrize("kind,cfg,case,seed", [("sca", mc.SCA_CFG, mc.sca_case, 10),
                                                ("tsa", mc.TSA_CFG, mc.tsa_case, 11),
                                                ("pred", mc.PRED_CFG, mc.pred_case, 12)])
def test_module_on_gpu_matches_reference_class(cuda, kind, cfg, case, seed):
    g = np.load(GOLD)
    m = build_attention(cfg)
    m.load_state_dict(mc.seeded_state(m, seed))
    m.eval().to(cuda)
    out, gq, gkv = mc.run_module(m, kind, case(), device=cuda)
    np.testing.assert_allclose(out.cpu().numpy(), g[f"{kind}_out"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gq.cpu().numpy(), g[f"{kind}_gq"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(gkv[:, ::6].cpu().numpy(), g[f"{kind}_gkv_s6"], rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("kind,boxes,seed", [("custom", False, 13), ("custom_boxes", True, 14)])
def test_detection_decoder_attention_on_gpu(cuda, kind, boxes, seed):
    g = np.load(GOLD_CUSTOM)
    m = build_attention(mc.CUSTOM_CFG)
    m.load_state_dict(mc.seeded_state(m, seed))
    m.eval().to(cuda)
    out, gq, gkv = mc.run_module(m, kind, mc.custom_case(boxes=boxes), device=cuda)
    np.testing.assert_allclose(out.cpu().numpy(), g[f"{kind}_out"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gq.cpu().numpy(), g[f"{kind}_gq"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(gkv.cpu().numpy(), g[f"{kind}_gkv"], rtol=1e-4, atol=5e-5)


def test_sca_fused_epilogue_equals_materialised_path(cuda):
    """SpatialCrossAttention through MSDeformableAttention3D with the softmax / sampling-location
    arithmetic inside the kernel (default) vs materialised `sampling_locations` / `attention_weights`."""
    m = build_attention(mc.SCA_CFG)
    m.load_state_dict(mc.seeded_state(m, 10))
    m.eval().to(cuda)
    assert m.deformable_attention.fuse_epilogue
    res = []
    for fused in (True, False):
        m.deformable_attention.fuse_epilogue = fused
        m.zero_grad(set_to_none=True)
        out, gq, gkv = mc.run_module(m, "sca", mc.sca_case(), device=cuda)
        res.append((out, gq, gkv, m.deformable_attention.sampling_offsets.weight.grad.clone(),
                    m.deformable_attention.attention_weights.weight.grad.clone()))
    for a, b, w in zip(res[0], res[1], ("out", "grad query", "grad key/value", "grad sampling_offsets.weight",
                                        "grad attention_weights.weight")):
        err = (a - b).abs().max().item()
        assert err <= 2e-5 * b.abs().max().item() + 1e-7, f"{w}: {err:.3e} vs {b.abs().max().item():.3e}"
