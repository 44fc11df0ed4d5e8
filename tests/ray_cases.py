"""Seeded inputs for the head ray sampler goldens/tests (shared with tools/make_golden_ray_head.py)."""
import torch

GRID = (4, 12, 10)        # Z, Y, X
NUM_WAY, STEP = 24, 1.0
FRAMES, BS, INTER = 2, 1, 2
LOSS_W = [[0.5, 1.0], [1.5, 2.0]]    # [frame][level]


def case(seed=0, M=90):
    g = torch.Generator().manual_seed(seed)
    Z, Y, X = GRID
    sigma = [torch.randn(BS, FRAMES, Z, Y, X, generator=g) for _ in range(INTER)]
    origin = torch.tensor([X / 2, Y / 2, Z / 2]) + 0.3 * torch.randn(BS, FRAMES, 3, generator=g)
    lo = torch.tensor([-2.0, -2.0, -1.0])
    hi = torch.tensor([X + 2.0, Y + 2.0, Z + 1.0])
    gt = lo + (hi - lo) * torch.rand(BS, M, 3, generator=g)          # some GT points outside
    tindex = torch.randint(0, FRAMES, (BS, M), generator=g).float()
    tindex[:, -6:] = -1                                              # padded rays
    gt[:, -6:] = float("nan")
    return dict(sigma=sigma, origin=origin, gt=gt, tindex=tindex)
