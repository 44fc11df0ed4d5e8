"""Parity at BASELINE.json sizes: the CUDA MSDA op against the fp64 CPU oracle on
  * the dense SpatialCrossAttention shape (6 cameras x 40000 queries, 4 levels x 8 points): ONE WHOLE
    camera -- output, grad_value, grad_sampling_loc, grad_attn_weight, all 40000 queries -- and a strided
    query slice of every other camera (per-query outputs; grad_value needs the whole camera);
  * the "rebatched" shape (6 x 10240 queries);
  * the TemporalSelfAttention shape (B=2, one 200x200 level, 4 points, Q=40000: the small-L*P kernels);
plus one direct oracle comparison for each fused entry point whose other tests are CUDA-vs-CUDA
(`vidar_msda_sca_*`, `vidar_ray_gumbel_*`).  Tolerances: tests/parity.py (norm + element-wise)."""
import numpy as np
import pytest
import torch

from oracle import msda_ref, ray_head_ref
from tests import parity
from vidar_b200 import msda, ray_head, synthetic

pytestmark = pytest.mark.gpu


def _gpu_all(d, cuda):
    g = {k: v.to(cuda) for k, v in d.items()}
    out = msda.ext_module.ms_deform_attn_forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"], im2col_step=64)
    gv = torch.zeros_like(g["value"])
    gl = torch.full_like(g["loc"], float("nan"))       # must be fully overwritten
    ga = torch.full_like(g["attn"], float("nan"))
    msda.ext_module.ms_deform_attn_backward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"], g["grad_out"],
                                            gv, gl, ga, im2col_step=64)
    torch.cuda.synchronize()
    return out.cpu(), gv.cpu(), gl.cpu(), ga.cpu()


def _cam(d, c, qsel=slice(None)):
    return dict(value=d["value"][c:c + 1], shapes=d["shapes"], lsi=d["lsi"], loc=d["loc"][c:c + 1, qsel].contiguous(),
                attn=d["attn"][c:c + 1, qsel].contiguous(), grad_out=d["grad_out"][c:c + 1, qsel].contiguous())


def _check_camera(res, d, c, qsel=None):
    """Camera c of the GPU result `res` against the oracle; qsel=None: every query, incl. grad_value."""
    out, gv, gl, ga = res
    sub = _cam(d, c, slice(None) if qsel is None else qsel)
    rout = msda_ref.msda_grid_sample(sub["value"].double(), sub["shapes"], sub["loc"].double(), sub["attn"].double())
    rgv, rgl, rga = msda_ref.msda_grid_sample_backward(sub["value"], sub["shapes"], sub["loc"], sub["attn"], sub["grad_out"])
    q = slice(None) if qsel is None else qsel
    tag = f"cam {c}" + ("" if qsel is None else " (query slice)")
    parity.close(out[c:c + 1, q], rout, f"{tag} output")
    keep = parity.off_kink(sub["loc"], sub["shapes"])
    assert keep.float().mean() > 0.999
    parity.close(gl[c:c + 1, q], rgl, f"{tag} grad_sampling_loc", keep=keep.unsqueeze(-1).expand_as(rgl))
    parity.close(ga[c:c + 1, q], rga, f"{tag} grad_attn_weight")
    if qsel is None:
        parity.close(gv[c:c + 1], rgv, f"{tag} grad_value")


@pytest.mark.parametrize("rows,whole_cam", [(None, 3), (10240, 1)], ids=["dense-40000", "rebatched-10240"])
def test_cfg2_forward_backward_full_size(cuda, rows, whole_cam):
    cpu = torch.device("cpu")
    d = synthetic.sca_like_inputs(cpu, cams=6, Q=40000, seed=0, rows=rows)
    res = _gpu_all(d, cuda)
    assert torch.isfinite(res[2]).all() and torch.isfinite(res[3]).all(), "grad_loc / grad_attn not fully written"
    _check_camera(res, d, whole_cam)                       # every query of one camera, all four outputs
    Q = d["loc"].shape[1]
    sl = slice(5, Q, 97)                                   # ~1% of the queries of every other camera
    for c in range(6):
        if c != whole_cam:
            _check_camera(res, d, c, sl)
    # the other cameras' grad_value: total mass per (camera, head) against the forward on an all-ones map
    g = {k: v.to(cuda) for k, v in d.items()}
    ones_out = msda.ext_module.ms_deform_attn_forward(torch.ones_like(g["value"]), g["shapes"], g["lsi"], g["loc"],
                                                      g["attn"], im2col_step=64)
    gv1 = torch.zeros_like(g["value"])
    gl1, ga1 = torch.empty_like(g["loc"]), torch.empty_like(g["attn"])
    msda.ext_module.ms_deform_attn_backward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"],
                                            torch.ones_like(g["grad_out"]), gv1, gl1, ga1, im2col_step=64)
    mass = gv1.double().sum((1, 3))                         # [cams, heads]   (x 32 channels)
    ref_mass = ones_out.view(6, Q, 8, 32).double().sum((1, 3))
    torch.testing.assert_close(mass, ref_mass, rtol=1e-6, atol=1e-3)


def test_tsa_shape_full_size(cuda):
    """TemporalSelfAttention's call: value [2, 40000, 8, 32] (prev / current BEV), one 200x200 level,
    4 points, 40000 queries -- the small-L*P kernels at their real size, every output, every query."""
    g = torch.Generator().manual_seed(21)
    B, Q, H, C, P = 2, 40000, 8, 32, 4
    levels = ((200, 200),)
    shapes, lsi = synthetic.level_tensors(levels)
    value = torch.randn(B, 40000, H, C, generator=g)
    iy, ix = torch.meshgrid(torch.arange(200), torch.arange(200), indexing="ij")
    ref = torch.stack([(ix.reshape(-1) + 0.5) / 200, (iy.reshape(-1) + 0.5) / 200], -1)       # ref_2d of the encoder
    loc = ref.view(1, Q, 1, 1, 1, 2) + 3.0 * torch.randn(B, Q, H, 1, P, 2, generator=g) / 200.0
    attn = torch.softmax(torch.randn(B, Q, H, P, generator=g), -1).view(B, Q, H, 1, P)
    d = dict(value=value, shapes=shapes, lsi=lsi, loc=loc.contiguous(), attn=attn.contiguous(),
             grad_out=torch.randn(B, Q, H * C, generator=g))
    res = _gpu_all(d, cuda)
    for b in range(B):
        _check_camera(res, d, b)


@pytest.mark.parametrize("Q", [4000])
def test_fused_sca_entry_points_against_oracle(cuda, Q):
    """vidar_msda_sca_forward/backward (softmax + sampling-location prologue inside the kernel) straight
    against the CPU oracle: the reference statements spatial_cross_attention.py:339-371 in fp64 torch
    followed by the grid_sample formulation, gradients by autograd."""
    g = torch.Generator().manual_seed(77)
    B, H, C, P, D = 2, 8, 32, 8, 4
    levels = synthetic.SCA_LEVELS
    L = len(levels)
    shapes, lsi = synthetic.level_tensors(levels)
    K = int((shapes[:, 0] * shapes[:, 1]).sum())
    value = torch.randn(B, K, H, C, generator=g)
    refp = torch.rand(B, Q, D, 2, generator=g) * 1.1 - 0.05
    offsets = 3.0 * torch.randn(B, Q, H, L, P, 2, generator=g)
    logits = 2.0 * torch.randn(B, Q, H, L * P, generator=g)
    grad = torch.randn(B, Q, H * C, generator=g)

    v1, o1, l1 = (t.to(cuda).requires_grad_(True) for t in (value, offsets, logits))
    out = msda.MSDeformAttn3DFusedFunction.apply(v1, shapes.to(cuda), lsi.to(cuda), refp.to(cuda), o1, l1)
    out.backward(grad.to(cuda))

    # the module's own fp32 statements build the locations (the kernel reproduces that division + add bit
    # for bit); from there on the oracle is fp64.  d(loc)/d(offsets) = 1 / (W_l, H_l).
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
    loc32 = ((offsets / wh[None, None, None, :, None, :]).view(B, Q, H, L, P // D, D, 2)
             + refp[:, :, None, None, None]).view(B, Q, H, L, P, 2)
    v2, l2 = value.double().requires_grad_(True), logits.double().requires_grad_(True)
    loc = loc32.double().requires_grad_(True)
    rout = msda_ref.msda_grid_sample(v2, shapes, loc, l2.softmax(-1).view(B, Q, H, L, P))
    rout.backward(grad.double())
    rgo = loc.grad / wh.double()[None, None, None, :, None, :]

    parity.close(out, rout, "sca_forward output")
    parity.close(v1.grad, v2.grad, "sca_backward grad_value")
    parity.close(l1.grad, l2.grad, "sca_backward grad_logits")
    keep = parity.off_kink(loc32, shapes)
    parity.close(o1.grad, rgo, "sca_backward grad_offsets", keep=keep.unsqueeze(-1).expand(-1, -1, -1, -1, -1, 2))


def test_ray_gumbel_entry_points_against_oracle(cuda):
    """vidar_ray_gumbel_forward/backward against the CPU oracle: oracle/ray_head_ref.sample_frame (the
    reference's _get_grid_features statements, F.grid_sample) + the reference's
    _custom_gumbel_softmax_distance statements (vidar_head_base.py:754-773) with the same noise."""
    gen = torch.Generator().manual_seed(9)
    Fr, Z, Y, X, R, K = 2, 8, 36, 40, 1500, 64
    sigma = torch.randn(Fr, Z, Y, X, generator=gen)
    origin = torch.tensor([X / 2, Y / 2, Z / 2]) + torch.randn(Fr, 3, generator=gen)
    pts = torch.rand(R, 3, generator=gen) * torch.tensor([X + 4.0, Y + 4.0, Z + 2.0]) - torch.tensor([2.0, 2.0, 1.0])
    frame = torch.randint(0, Fr, (R,), generator=gen).sort()[0].to(torch.int32)
    noise = -torch.empty(R, K).exponential_(generator=gen).log()
    gout = torch.randn(R, generator=gen)

    s1 = sigma.to(cuda).requires_grad_(True)
    d1 = ray_head.gumbel_distance(s1, origin.to(cuda), pts.to(cuda), frame.to(cuda), K, 1.0, noise.to(cuda))
    (d1 * gout.to(cuda)).sum().backward()

    s2 = sigma.clone().requires_grad_(True)
    d2 = torch.zeros(R)
    for f in range(Fr):
        sel = torch.nonzero(frame == f).squeeze(-1)
        logits, length, valid = ray_head_ref.sample_frame(s2[f], origin[f], pts[sel], K, 1.0, with_gt=True)
        lg, ln = logits[valid][:, 1:], length[valid][:, 1:]              # the dense term has no GT slot
        # vidar_head_base.py:757-758 with F.gumbel_softmax(hard=True) written out on the given noise
        y_soft = (lg + noise[sel][valid]).softmax(-1)
        y_hard = torch.zeros_like(lg).scatter_(-1, y_soft.max(-1, keepdim=True)[1], 1.0)
        pred = ((y_hard - y_soft.detach() + y_soft) * ln).sum(-1).detach()
        # :761-772
        e = torch.exp(lg - lg.max(-1, keepdim=True)[0])
        prob_next = (e * (ln > pred.unsqueeze(-1)).float()).sum(-1) / e.sum(-1)
        d2[sel[valid]] = (1 - prob_next.detach() + prob_next) * pred
    (d2 * gout).sum().backward()
    np.testing.assert_allclose(d1.detach().cpu().numpy(), d2.detach().numpy(), rtol=1e-5, atol=1e-5)
    parity.close(s1.grad, s2.grad, "ray_gumbel grad_sigma", elementwise=False)
