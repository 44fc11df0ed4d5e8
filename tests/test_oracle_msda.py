"""Pin the MSDA oracle: grid_sample statement == per-corner statement == the same-lineage
implementation shipped in `transformers` (mmcv itself is not in this image)."""
import numpy as np
import pytest
import torch

from oracle import msda_ref
from tests.inputs import msda_inputs

LEVELS = ((9, 13), (5, 7), (3, 4))


def _case(mode, seed):
    return msda_inputs(B=2, Q=37, H=3, C=8, levels=LEVELS, P=4, seed=seed, mode=mode,
                       dtype=torch.float64)


@pytest.mark.parametrize("mode", ["local", "stress"])
def test_grid_sample_equals_explicit_forward_backward(mode):
    d = _case(mode, 1)
    out_gs = msda_ref.msda_grid_sample(d["value"], d["shapes"], d["loc"], d["attn"])
    gv, gl, ga = msda_ref.msda_grid_sample_backward(d["value"], d["shapes"], d["loc"], d["attn"],
                                                    d["grad_out"])
    out, ev, el, ea = msda_ref.msda_explicit(d["value"].numpy(), d["shapes"].numpy(), d["lsi"].numpy(),
                                             d["loc"].numpy(), d["attn"].numpy(), d["grad_out"].numpy())
    np.testing.assert_allclose(out, out_gs.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ev, gv.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(el, gl.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ea, ga.numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("mode", ["local", "stress"])
def test_matches_transformers_deformable_detr(mode):
    tr = pytest.importorskip("transformers.models.deformable_detr.modeling_deformable_detr")
    d = _case(mode, 2)
    ref = tr.MultiScaleDeformableAttention().forward(
        d["value"], d["shapes"], [tuple(x) for x in d["shapes"].tolist()], d["lsi"], d["loc"],
        d["attn"], 64)
    out = msda_ref.msda_grid_sample(d["value"], d["shapes"], d["loc"], d["attn"])
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_out_of_range_samples_contribute_nothing():
    d = _case("local", 3)
    loc = d["loc"].clone()
    loc[:, :, :, 0] = -0.5          # whole level 0 far outside
    loc[:, :, :, 1, :, 0] = 1.5
    base = msda_ref.msda_grid_sample(d["value"], d["shapes"], loc, d["attn"])
    attn = d["attn"].clone()
    attn[:, :, :, 0] = 0
    attn[:, :, :, 1] = 0
    zeroed = msda_ref.msda_grid_sample(d["value"], d["shapes"], d["loc"], attn)
    torch.testing.assert_close(base, zeroed, rtol=1e-12, atol=1e-12)
