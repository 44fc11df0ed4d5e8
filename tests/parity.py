"""Shared comparison helpers of the parity tests.

Two statements of north_star's "1e-4 relative fp32":
  * norms      max|a-b| <= 1e-4 max|b|   and   ||a-b||_2 <= 2e-5 ||b||_2
  * elements   |a-b| <= 1e-4 |b| + floor,  floor = 1e-5 max|b|
The floor is what fp32 *input* quantisation leaves on an element that is a cancelling sum (a
pixel coordinate near 100 has an ulp of 7.6e-6 px); it is ten times tighter than the max-norm
bound, so a regression confined to small-magnitude outputs is caught.
"""
import torch

RTOL, FLOOR = 1e-4, 1e-5


def to64(t):
    return torch.as_tensor(t).detach().cpu().to(torch.float64)


def close(a, b, what, keep=None, elementwise=True):
    a, b = to64(a), to64(b)
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if keep is not None:
        a, b = a[keep], b[keep]
    if b.numel() == 0:
        return
    err = (a - b).abs()
    bmax, bl2 = b.abs().max().item() + 1e-30, b.norm().item() + 1e-30
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    assert err.max().item() <= RTOL * bmax, f"{what}: max err {err.max().item():.3e} vs max|ref| {bmax:.3e}"
    assert err.norm().item() <= 2e-5 * bl2, f"{what}: L2 err {err.norm().item():.3e} vs ||ref|| {bl2:.3e}"
    if elementwise:
        bound = RTOL * b.abs() + FLOOR * bmax
        bad = err > bound
        if bad.any():
            i = int(torch.argmax((err - bound).flatten()))
            raise AssertionError(
                f"{what}: {int(bad.sum())} of {b.numel()} elements outside |a-b| <= 1e-4|b| + 1e-5 max|b|; worst "
                f"a={a.flatten()[i].item():.6e} b={b.flatten()[i].item():.6e} (max|b| {bmax:.3e})")


def off_kink(loc, shapes, eps=1e-4):
    """[B,Q,H,L,P] bool: sample farther than eps px from every pixel-centre line (where
    d(out)/d(loc) is discontinuous and either one-sided gradient is legitimate)."""
    loc = to64(loc)
    shapes = torch.as_tensor(shapes).cpu()
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).double()      # (W, H) per level
    px = loc * wh.view(1, 1, 1, -1, 1, 2) - 0.5
    frac = (px - px.round()).abs()
    return (frac > eps).all(-1)
