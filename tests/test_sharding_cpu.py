"""Multi-rank host logic over gloo on CPU (world_size 2 and 3): row/ray partitioning, the
all-gather of BEV rows and the grad reductions reproduce the single-process result.  The
per-rank compute is the CPU oracle here (the CUDA kernels are covered by the gpu tests)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vidar_b200 import sharding


def test_partitions_cover_everything_once():
    for world in (1, 2, 3, 4, 8):
        rows = []
        for r in range(world):
            for c, q0, q1 in sharding.shard_rows(r, world, 6, 40000):
                rows += [(c, q0, q1)]
        total = sum(q1 - q0 for _, q0, q1 in rows)
        assert total == 240000
        sizes = [sum(q1 - q0 for _, q0, q1 in sharding.shard_rows(r, world, 6, 40000)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
        lo = [sharding.shard_range(30000, r, world) for r in range(world)]
        assert lo[0][0] == 0 and lo[-1][1] == 30000 and all(a[1] == b[0] for a, b in zip(lo, lo[1:]))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dvr_ref, msda_ref
    from tests.inputs import dvr_inputs_lidar, msda_inputs
    cams, Q, H, C, P = 3, 50, 2, 8, 2
    levels = ((6, 8), (3, 4))
    d = msda_inputs(cams, Q, H, C, levels, P, seed=5, dtype=torch.float64)
    groups = sharding.camera_groups(world, cams, Q, rank)
    outs, gvals = [], {}
    for c, q0, q1 in sharding.shard_rows(rank, world, cams, Q):
        v = d["value"][c:c + 1].clone().requires_grad_(True)
        o = msda_ref.msda_grid_sample(v, d["shapes"], d["loc"][c:c + 1, q0:q1], d["attn"][c:c + 1, q0:q1])
        o.backward(d["grad_out"][c:c + 1, q0:q1])
        outs.append(o.detach().view(-1, H * C))
        gvals[c] = gvals.get(c, 0) + v.grad
    rows = sharding.gather_rows(torch.cat(outs, 0), world, cams * Q)
    for c, g in groups.items():
        dist.all_reduce(gvals[c], group=g)
    # rays: shard, render, all-reduce grad_sigma
    sigma, origin, points, tindex = dvr_inputs_lidar(M=400, T=2, grid=(4, 24, 24), seed=2)
    lo, hi = sharding.shard_range(400, rank, world)
    pred, gt, grad = dvr_ref.render(sigma, origin, points[:, lo:hi], tindex[:, lo:hi], "l2")
    g = torch.from_numpy(grad).double()
    dist.all_reduce(g)
    if rank == 0:
        torch.save(dict(rows=rows, gvals=gvals, grad_sigma=g), tmp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_single_process(tmp_path, world):
    from oracle import dvr_ref, msda_ref
    from tests.inputs import dvr_inputs_lidar, msda_inputs
    port = 29500 + os.getpid() % 500 + world
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    cams, Q, H, C, P = 3, 50, 2, 8, 2
    d = msda_inputs(cams, Q, H, C, ((6, 8), (3, 4)), P, seed=5, dtype=torch.float64)
    v = d["value"].clone().requires_grad_(True)
    full = msda_ref.msda_grid_sample(v, d["shapes"], d["loc"], d["attn"])
    full.backward(d["grad_out"])
    torch.testing.assert_close(got["rows"], full.detach().view(-1, H * C), rtol=1e-12, atol=1e-12)
    for c, g in got["gvals"].items():           # rank 0's cameras, summed over the ranks sharing them
        torch.testing.assert_close(g[0], v.grad[c], rtol=1e-10, atol=1e-12)
    sigma, origin, points, tindex = dvr_inputs_lidar(M=400, T=2, grid=(4, 24, 24), seed=2)
    _, _, grad = dvr_ref.render(sigma, origin, points, tindex, "l2")
    np.testing.assert_allclose(got["grad_sigma"].numpy(), grad, rtol=1e-5, atol=1e-6)


def _rows_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)
    R, C = 23, 5                                    # 23 rows: uneven blocks, the last one short
    parts = [torch.randn(R, C, generator=g, dtype=torch.float64) for _ in range(world)]
    w = torch.randn(C, C, generator=g, dtype=torch.float64)
    gout = torch.randn(R, C, generator=g, dtype=torch.float64)
    lin = torch.nn.Linear(C, C, bias=False).double()
    with torch.no_grad():
        lin.weight.copy_(w)
    sharding.mark_partial(lin)
    x = parts[rank].clone().requires_grad_(True)           # this rank's partial sum
    rep = torch.ones(R, C, dtype=torch.float64, requires_grad=True)     # a replicated tensor consumed per rank
    rows = sharding.reduce_scatter_rows(x * sharding.sum_grad(rep, dist.group.WORLD), dist.group.WORLD)
    lo, hi = sharding.row_range(R, rank, world)
    assert rows.shape[0] == hi - lo
    full = sharding.all_gather_rows(lin(rows), R, dist.group.WORLD)
    (full * gout).sum().backward()
    n = sharding.allreduce_partial_grads(lin, dist.group.WORLD)
    assert n == C * C
    torch.save(dict(full=full.detach(), gx=x.grad, gw=lin.weight.grad, grep=rep.grad), os.path.join(tmp, f"rows{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_collectives_with_autograd(tmp_path, world):
    """reduce_scatter_rows -> row-wise layer -> all_gather_rows equals the single-process computation:
    output, gradient of every rank's partial sum, the summed gradient of a replicated input (sum_grad)
    and -- after allreduce_partial_grads -- the row-wise layer's weight gradient."""
    port = 29700 + os.getpid() % 200 + world
    mp.spawn(_rows_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(3)
    R, C = 23, 5
    parts = [torch.randn(R, C, generator=g, dtype=torch.float64) for _ in range(world)]
    w = torch.randn(C, C, generator=g, dtype=torch.float64).requires_grad_(True)
    gout = torch.randn(R, C, generator=g, dtype=torch.float64)
    xs = [p.clone().requires_grad_(True) for p in parts]
    rep = torch.ones(R, C, dtype=torch.float64, requires_grad=True)
    full = (sum(x * rep for x in xs)) @ w.t()
    (full * gout).sum().backward()
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rows{r}.pt"), weights_only=False)
        torch.testing.assert_close(got["full"], full.detach())
        torch.testing.assert_close(got["gx"], xs[r].grad)
        torch.testing.assert_close(got["gw"], w.grad)
        torch.testing.assert_close(got["grep"], rep.grad)


def _partial_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lin_a, lin_b = torch.nn.Linear(3, 2).double(), torch.nn.Linear(3, 2).double()
    with torch.no_grad():
        for m in (lin_a, lin_b):
            m.weight.fill_(0.5)
            m.bias.zero_()
    model = torch.nn.ModuleList([lin_a, lin_b])
    sharding.mark_partial(model)
    x = torch.full((4, 3), float(rank + 1), dtype=torch.float64)
    # rank 0 uses both layers, every other rank only the second one (a rank that owns no camera never runs the backbone)
    y = lin_b(x).sum() + (lin_a(x).sum() if rank == 0 else 0.0)
    y.backward()
    assert (lin_a.weight.grad is None) == (rank != 0)
    sharding.allreduce_partial_grads(model, dist.group.WORLD)
    torch.save(dict(ga=lin_a.weight.grad, gb=lin_b.weight.grad), os.path.join(tmp, f"pg{rank}.pt"))
    dist.destroy_process_group()


def test_partial_grad_bucket_is_identical_on_ranks_without_a_gradient(tmp_path):
    """A tagged parameter that one rank never used (grad None) still takes part in the bucketed all-reduce
    (that rank sends zeros) and ends up with the summed gradient everywhere."""
    world = 3
    mp.spawn(_partial_worker, args=(world, 29900 + os.getpid() % 90, str(tmp_path)), nprocs=world, join=True)
    ga = torch.full((2, 3), 4.0 * 1, dtype=torch.float64)                      # only rank 0: sum over 4 rows of x = 1
    gb = torch.full((2, 3), 4.0 * (1 + 2 + 3), dtype=torch.float64)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"pg{r}.pt"), weights_only=False)
        torch.testing.assert_close(got["ga"], ga)
        torch.testing.assert_close(got["gb"], gb)
