"""NN / Chamfer: oracle pinned to the reference's own CPU build (golden), CUDA op vs both."""
import os

import numpy as np
import pytest
import torch

from oracle import knn_ref
from tests.knn_cases import case

GOLD = os.path.join(os.path.dirname(__file__), "golden", "knn.npz")


def test_oracle_matches_reference_cpu_build():
    g = np.load(GOLD)
    c = case()
    d, i = knn_ref.nn(c["a"].numpy(), c["b"].numpy(), c["la"].numpy(), c["lb"].numpy())
    np.testing.assert_array_equal(i, g["idx"][..., 0])
    np.testing.assert_array_equal(d, g["dists"][..., 0])                # same fp32 operation order
    ga, gb = knn_ref.nn_backward(c["a"].numpy(), c["b"].numpy(), i, c["g"].numpy()[..., 0], c["la"].numpy())
    np.testing.assert_allclose(ga, g["ga"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gb, g["gb"], rtol=1e-5, atol=1e-5)
    f, b = knn_ref.chamfer(c["a"].numpy(), c["b"].numpy(), bidirectional=True, reduction="sum")
    np.testing.assert_allclose([f, b], [g["cham_fwd"], g["cham_bwd"]], rtol=1e-5)


@pytest.mark.gpu
def test_cuda_knn_and_chamfer_match_reference(cuda):
    from vidar_b200 import chamfer
    g = np.load(GOLD)
    c = case()
    a = c["a"].to(cuda).requires_grad_(True)
    b = c["b"].to(cuda).requires_grad_(True)
    k = chamfer.knn_points(a, b, lengths1=c["la"].to(cuda), lengths2=c["lb"].to(cuda), K=1)
    np.testing.assert_array_equal(k.idx.cpu().numpy(), g["idx"])
    np.testing.assert_array_equal(k.dists.detach().cpu().numpy(), g["dists"])
    (k.dists * c["g"].to(cuda)).sum().backward()
    np.testing.assert_allclose(a.grad.cpu().numpy(), g["ga"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(b.grad.cpu().numpy(), g["gb"], rtol=1e-5, atol=1e-5)
    cd = chamfer.ChamferDistance()
    f, bw, info = cd(c["a"].to(cuda), c["b"].to(cuda), bidirectional=True, reduction="sum")
    np.testing.assert_allclose([f.item(), bw.item()], [g["cham_fwd"], g["cham_bwd"]], rtol=1e-5)
    np.testing.assert_array_equal(info[1].cpu().numpy(), g["info_fi"])
    np.testing.assert_array_equal(info[3].cpu().numpy(), g["info_bi"])
    m, _ = cd(c["a"].to(cuda), c["b"].to(cuda), reduction="mean")
    r, _ = cd(c["a"].to(cuda), c["b"].to(cuda), reverse=True, reduction=None)
    np.testing.assert_allclose(m.item(), g["cham_mean"], rtol=1e-5)
    np.testing.assert_allclose(r.cpu().numpy(), g["cham_rev"], rtol=1e-5)
    with pytest.raises(NotImplementedError):
        chamfer.knn_points(a, b, K=3)


@pytest.mark.gpu
def test_cuda_knn_full_size_properties(cuda):
    """30k x 30k points (ViDAR eval size): NN of a cloud against itself is the identity with
    distance 0; a subset agrees with the oracle; chamfer(a, a) == 0."""
    from vidar_b200 import chamfer
    g = torch.Generator().manual_seed(1)
    a = (torch.rand(1, 30000, 3, generator=g) * 100).to(cuda)
    b = (torch.rand(1, 29000, 3, generator=g) * 100).to(cuda)
    k = chamfer.knn_points(a, a)
    assert torch.equal(k.idx[0, :, 0], torch.arange(30000, device=cuda)) and float(k.dists.max()) == 0.0
    k2 = chamfer.knn_points(a, b)
    d, i = knn_ref.nn(a[:, ::97].cpu().numpy(), b.cpu().numpy())
    np.testing.assert_array_equal(k2.idx[0, ::97, 0].cpu().numpy(), i[0])
    np.testing.assert_array_equal(k2.dists[0, ::97, 0].cpu().numpy(), d[0])
    assert float(chamfer.compute_chamfer_distance(a[0], a[0])) == 0.0
