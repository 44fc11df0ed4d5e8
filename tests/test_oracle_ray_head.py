"""Pin oracle/ray_head_ref.py against the reference's own methods (tests/golden/ray_head.npz,
made by tools/make_golden_ray_head.py)."""
import os

import numpy as np
import torch

from oracle import ray_head_ref as ref
from tests import ray_cases as rc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ray_head.npz")


def test_grid_features_ce_and_gradients_match_reference():
    g = np.load(GOLD)
    c = rc.case()
    sig = [s.clone().requires_grad_(True) for s in c["sigma"]]
    r_mask, r_feat, r_w, r_len = ref.grid_features(c["origin"], c["gt"], c["tindex"], sig, rc.LOSS_W,
                                                   rc.STEP, rc.NUM_WAY)
    np.testing.assert_array_equal(r_mask.numpy(), g["r_mask"])
    np.testing.assert_allclose(r_feat.detach().numpy(), g["r_feat"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(r_w.numpy(), g["r_w"])
    np.testing.assert_allclose(r_len.numpy(), g["r_len"], rtol=1e-6)
    loss = ref.ce_loss(r_feat, r_w)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-6)
    loss.backward()
    np.testing.assert_allclose(sig[0].grad.numpy(), g["grad_sigma0"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(sig[1].grad.numpy(), g["grad_sigma1"], rtol=1e-5, atol=1e-7)


def test_decode_matches_reference():
    g = np.load(GOLD)
    c = rc.case()
    sigma = c["sigma"][-1]
    for b in range(rc.BS):
        for f in range(rc.FRAMES):
            sel = c["tindex"][b] == f
            pred, idx = ref.decode_frame(sigma[b, f], c["origin"][b, f], c["gt"][b][sel], rc.NUM_WAY, rc.STEP)
            np.testing.assert_array_equal(idx.numpy(), g["decode_idx"][b][sel.numpy()])
            np.testing.assert_allclose(pred.numpy(), g["decode_pred"][b][sel.numpy()], rtol=1e-6)
