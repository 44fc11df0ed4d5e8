"""numpy front-end of oracle/dvr_ref.c (CPU restatement of dvr / dvxlr / dvxlr_v2).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU
legs.  Function names and return lists mirror the reference bindings
(third_lib/dvr/dvr.cpp:65-69, third_lib/dvxlr/dvxlr.cpp:61-65, dvxlr_v2.cpp:67-70) with numpy
arrays in place of CUDA tensors.
"""
import ctypes as C

import numpy as np

from . import build as _build

MAX_D = 1026  # third_lib/dvxlr/dvxlr.cu:10
_LOSS = {"l1": 0, "bce": 0, "l2": 1, "absrel": 2}   # dvr.cu:661-672
_PHASE = {"test": 0, "train": 1}                    # dvr.cu:358-365

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dims(sigma, origin, points):
    N, T, Z, Y, X = sigma.shape
    To = origin.shape[1]
    M = points.shape[1]
    return [C.c_int(v) for v in (N, M, T, To, Z, Y, X)]


def init(points, tindex, grid):
    points, tindex = _f(points), _f(tindex)
    T, Z, Y, X = grid
    N, M = points.shape[:2]
    occ = np.zeros((N, T, Z, Y, X), np.float32)
    lib().oracle_dvr_init(_p(points), _p(tindex), _p(occ), N, M, T, Z, Y, X)
    return occ


def render_forward(sigma, origin, points, tindex, grid=None, phase_name="test"):
    sigma, origin, points, tindex = map(_f, (sigma, origin, points, tindex))
    if phase_name not in _PHASE:
        raise ValueError(f"UNKNOWN PHASE NAME: {phase_name}")
    N, M = points.shape[:2]
    pred = -np.ones((N, M), np.float32)
    gt = -np.ones((N, M), np.float32)
    lib().oracle_dvr_render_forward(_p(sigma), _p(origin), _p(points), _p(tindex), _p(pred), _p(gt),
                                    *_dims(sigma, origin, points), C.c_int(_PHASE[phase_name]))
    return [pred, gt]


def render(sigma, origin, points, tindex, loss_name="l1"):
    sigma, origin, points, tindex = map(_f, (sigma, origin, points, tindex))
    if loss_name not in _LOSS:
        raise ValueError(f"UNKNOWN LOSS TYPE: {loss_name}")
    N, M = points.shape[:2]
    pred = -np.ones((N, M), np.float32)
    gt = -np.ones((N, M), np.float32)
    grad = np.zeros_like(sigma)
    lib().oracle_dvr_render(_p(sigma), _p(origin), _p(points), _p(tindex), _p(pred), _p(gt), _p(grad),
                            *_dims(sigma, origin, points), C.c_int(_LOSS[loss_name]))
    return [pred, gt, grad]


def dvxlr_forward(sigma, origin, points, tindex):
    sigma, origin, points, tindex = map(_f, (sigma, origin, points, tindex))
    N, M = points.shape[:2]
    pred = -np.ones((N, M), np.float32)
    gt = -np.ones((N, M), np.float32)
    lib().oracle_dvxlr_forward(_p(sigma), _p(origin), _p(points), _p(tindex), _p(pred), _p(gt),
                               *_dims(sigma, origin, points))
    return [pred, gt]


def dvxlr_render(sigma, origin, points, tindex, sigma_regul=None, max_d=MAX_D):
    """dvxlr.render (sigma_regul None) -> [pred, gt, dd_dsigma, indices];
    dvxlr_v2.render_v2 -> [pred, gt, dd_dsigma, indices, ray_pred, indicator]."""
    sigma, origin, points, tindex = map(_f, (sigma, origin, points, tindex))
    N, M = points.shape[:2]
    pred = -np.ones((N, M), np.float32)
    gt = -np.ones((N, M), np.float32)
    dd = np.zeros((N, M, max_d), np.float32)
    idx = np.zeros((N, M, max_d, 3), np.float32)
    ray_pred = indicator = None
    if sigma_regul is not None:
        sigma_regul = _f(sigma_regul)
        ray_pred = np.zeros((N, M, max_d), np.float32)
        indicator = -np.ones((N, M, max_d), np.float32)
    lib().oracle_dvxlr_render(_p(sigma), _p(origin), _p(points), _p(tindex), _p(sigma_regul),
                              _p(pred), _p(gt), _p(dd), _p(idx), _p(ray_pred), _p(indicator),
                              *_dims(sigma, origin, points), C.c_int(max_d))
    if sigma_regul is None:
        return [pred, gt, dd, idx]
    return [pred, gt, dd, idx, ray_pred, indicator]


def dvxlr_get_grad_sigma(elementwise_mult, indices, tindex, sigma_like, indicator=None,
                         grad_ray_pred=None):
    em, indices, tindex = map(_f, (elementwise_mult, indices, tindex))
    N, T, Z, Y, X = sigma_like.shape
    M, max_d = em.shape[1], em.shape[2]
    g = np.zeros((N, T, Z, Y, X), np.float32)
    g2 = None
    if indicator is not None:
        indicator, grad_ray_pred = _f(indicator), _f(grad_ray_pred)
        g2 = np.zeros_like(g)
    lib().oracle_dvxlr_get_grad_sigma(_p(em), _p(indices), _p(tindex), _p(indicator),
                                      _p(grad_ray_pred), _p(g), _p(g2),
                                      C.c_int(N), C.c_int(M), C.c_int(T), C.c_int(Z), C.c_int(Y),
                                      C.c_int(X), C.c_int(max_d))
    return [g] if g2 is None else [g, g2]


def dvxlr_autograd_backward(sigma, origin, points, tindex, grad_pred, sigma_regul=None,
                            grad_ray_pred=None, max_d=MAX_D):
    """What DifferentiableVoxelRendering[V2].backward returns
    (e2e_predictor_utils.py:102-113 / :134-141): render -> gradpred*dd (NaN->0 for v1) ->
    get_grad_sigma."""
    out = dvxlr_render(sigma, origin, points, tindex, sigma_regul, max_d)
    dd, idx = out[2], out[3]
    em = np.asarray(grad_pred, np.float32)[..., None] * dd
    if sigma_regul is None:
        em[np.isnan(em)] = 0.0
        return dvxlr_get_grad_sigma(em, idx, tindex, sigma)
    return dvxlr_get_grad_sigma(em, idx, tindex, sigma, out[5], grad_ray_pred)
