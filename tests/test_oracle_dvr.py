"""Properties of the dvr/dvxlr CPU oracle (oracle/dvr_ref.c) that hold for the reference
algorithm independent of any implementation: known-answer rays, finite differences of the
analytic gradient, consistency between the list path and the fused loss path."""
import numpy as np
import pytest

from oracle import dvr_ref
from tests.inputs import dvr_inputs_cfg1, dvr_inputs_lidar


def test_known_answer_axis_ray():
    # 1 x 1 x 4 grid along x, origin in voxel 0 at x=0.5, ray to x=3.5 (inside voxel 3)
    sigma = np.array([0.5, 1.0, 2.0, 4.0], np.float32).reshape(1, 1, 1, 1, 4)
    origin = np.array([[[0.5, 0.5, 0.5]]], np.float32)
    points = np.array([[[3.5, 0.5, 0.5]]], np.float32)
    tindex = np.zeros((1, 1), np.float32)
    # dvr.render uses the standard boundary: crossings at d = 0.5, 1.5, 2.5, 3.5
    d = np.array([0.5, 1.5, 2.5, 3.5])
    dt = np.array([0.5, 1.0, 1.0, 1.0])
    csd = np.cumsum(np.array([0.5, 1.0, 2.0, 4.0]) * dt)
    T = np.exp(-csd)
    p = np.concatenate([[1 - T[0]], T[:-1] - T[1:]])
    expect = (p * d).sum() + T[-1] * d[-1]
    pred, gt, grad = dvr_ref.render(sigma, origin, points, tindex, "l2")
    assert pred[0, 0] == pytest.approx(expect, rel=1e-6)
    assert gt[0, 0] == pytest.approx(3.0, rel=1e-6)          # min(|p-o|, max_d=3.5)
    # analytic gradient of 0.5*(pred-gt)^2
    dd = np.array([-dt[j] * (T[j:-1] * np.diff(d)[j:]).sum() for j in range(4)])
    np.testing.assert_allclose(grad.ravel(), (expect - 3.0) * dd, rtol=1e-5, atol=1e-9)
    assert grad.ravel()[-1] == 0.0                           # last voxel never gets gradient


def test_padded_and_missing_rays_keep_minus_one():
    sigma, origin, points, tindex = dvr_inputs_cfg1(seed=0)
    pred, gt = dvr_ref.render_forward(sigma, origin, points, tindex, None, "test")
    assert (pred[0, -16:] == -1).all() and (gt[0, -16:] == -1).all()
    assert (pred[0, :-16] >= 0).all()      # origin is inside the grid -> every ray has a path


def test_zero_sigma_gives_exit_distance_and_test_phase_keeps_raw_gt():
    sigma, origin, points, tindex = dvr_inputs_cfg1(seed=1, integer_origin=False, pad=0)
    z = np.zeros_like(sigma)
    pred, gt_train = dvr_ref.render_forward(z, origin, points, tindex, None, "train")
    _, gt_test = dvr_ref.render_forward(z, origin, points, tindex, None, "test")
    raw = np.linalg.norm(points - origin[:, :1], axis=-1).astype(np.float32)
    np.testing.assert_allclose(gt_test, raw, rtol=1e-6)
    np.testing.assert_array_equal(gt_train, np.minimum(gt_test, pred))  # pred == max_d here
    assert (gt_train <= gt_test).all()


@pytest.mark.parametrize("integer_origin", [True, False])
@pytest.mark.parametrize("loss", ["l1", "l2", "absrel"])
def test_render_gradient_matches_finite_differences(integer_origin, loss):
    sigma, origin, points, tindex = dvr_inputs_cfg1(seed=2, integer_origin=integer_origin, M=200)
    pred, gt, grad = dvr_ref.render(sigma, origin, points, tindex, loss)
    valid = pred[0] >= 0

    def total(s):
        p, g, _ = dvr_ref.render(s, origin, points, tindex, loss)
        p, g = p[0][valid].astype(np.float64), g[0][valid].astype(np.float64)
        if loss == "l1":
            return np.abs(p - g).sum()
        if loss == "l2":
            return 0.5 * ((p - g) ** 2).sum()
        return (np.abs(p - g) / g).sum()

    rng = np.random.default_rng(0)
    flat = np.flatnonzero(np.abs(grad.ravel()) > 1e-3)
    for i in rng.choice(flat, 12, replace=False):
        eps = 1e-2
        sp, sm = sigma.copy().ravel(), sigma.copy().ravel()
        sp[i] += eps
        sm[i] -= eps
        fd = (total(sp.reshape(sigma.shape)) - total(sm.reshape(sigma.shape))) / (2 * eps)
        assert fd == pytest.approx(grad.ravel()[i], rel=2e-2, abs=2e-3)


def test_dvxlr_lists_reproduce_fused_gradient_and_v2_extras():
    sigma, origin, points, tindex = dvr_inputs_lidar(M=600, T=2, grid=(8, 40, 40), seed=3, pad=8)
    rng = np.random.default_rng(1)
    regul = rng.standard_normal(sigma.shape).astype(np.float32)
    pred, gt, dd, idx, ray_pred, ind = dvr_ref.dvxlr_render(sigma, origin, points, tindex, regul)
    p1, g1, dd1, idx1 = dvr_ref.dvxlr_render(sigma, origin, points, tindex)
    np.testing.assert_array_equal(pred, p1)
    np.testing.assert_array_equal(dd, dd1)
    fwd_p, fwd_g = dvr_ref.dvxlr_forward(sigma, origin, points, tindex)
    np.testing.assert_array_equal(fwd_p, pred)
    np.testing.assert_array_equal(fwd_g, gt)
    # indicator: -1 pad, 0 valid, at most one 1 per ray; ray_pred gathers sigma_regul
    n_valid = (ind >= 0).sum(-1)
    assert (ind == 1).sum(-1).max() <= 1
    assert ((ind[..., 1:] >= 0) <= (ind[..., :-1] >= 0)).all()      # valid slots are a prefix
    r, i = 5, 0
    z, y, x = idx[0, r, i].astype(int)
    assert ray_pred[0, r, i] == regul[0, int(tindex[0, r]), z, y, x]
    # consecutive recorded voxels of a ray differ (duplicate merge)
    for r in range(0, 500, 37):
        k = int(n_valid[0, r])
        v = idx[0, r, :k]
        assert k > 0 and (np.abs(np.diff(v, axis=0)).sum(-1) > 0).all()
    # scatter of gradpred*dd == directional derivative of sum(gradpred*pred)
    gp = rng.standard_normal(pred.shape).astype(np.float32)
    (gs,) = dvr_ref.dvxlr_autograd_backward(sigma, origin, points, tindex, gp)
    direction = rng.standard_normal(sigma.shape).astype(np.float32)
    eps = 1e-3
    pp, _ = dvr_ref.dvxlr_forward(sigma + eps * direction, origin, points, tindex)
    pm, _ = dvr_ref.dvxlr_forward(sigma - eps * direction, origin, points, tindex)
    ok = pred >= 0
    fd = ((pp - pm).astype(np.float64)[ok] * gp[ok]).sum() / (2 * eps)
    assert fd == pytest.approx((gs.astype(np.float64) * direction).sum(), rel=5e-3)


def test_init_marks_endpoint_voxels():
    sigma, origin, points, tindex = dvr_inputs_cfg1(seed=4)
    occ = dvr_ref.init(points, tindex, [1, 8, 50, 50])
    # int() truncates toward zero (dvr.cu:52-54): coordinates in (-1, 0) land in voxel 0
    v = np.trunc(np.nan_to_num(points[0], nan=-9.0)).astype(int)
    inside = ((v >= 0).all(-1) & (v[:, 0] < 50) & (v[:, 1] < 50) & (v[:, 2] < 8) & (tindex[0] >= 0))
    vox = np.unique(v[inside], axis=0)
    assert (points[0][inside] < 0).any()          # the quirk is exercised
    assert occ.sum() == len(vox)
