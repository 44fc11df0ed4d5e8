"""2-GPU checks (skipped on a 1-GPU box): the cell-sharded LatentRendering core equals the
single-GPU op.  Launched through torch.multiprocessing with NCCL on 127.0.0.1."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vidar_b200.modules.latent_rendering import latent_render_core
    g = torch.Generator().manual_seed(0)
    occ = torch.randn(1, 60, 50, 16, generator=g).cuda()
    feat = torch.randn(1, 60, 50, 16, generator=g).cuda()
    gp = torch.randn(1, 60, 50, 16, generator=g).cuda()
    gq = torch.randn(1, 3000, 16, generator=g).cuda()
    res = []
    for group in (None, dist.group.WORLD):
        o, f = occ.clone().requires_grad_(True), feat.clone().requires_grad_(True)
        p, q = latent_render_core(o, f, 64, 0.5, 1e-3, 1, group)
        ((p * gp).sum() + (q * gq).sum()).backward()
        res.append([t.detach().cpu() for t in (p, q, o.grad, f.grad)])
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_latent_core_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, 29640 + os.getpid() % 200, out), nprocs=2, join=True)
    single, sharded = torch.load(out, weights_only=False)
    for a, b, w in zip(sharded, single, ["prob", "pooled", "grad_occ", "grad_feat"]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()), msg=lambda m: f"{w}: {m}")


def _module_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tests import latent_cases as lc
    import vidar_b200.modules  # noqa: F401
    from vidar_b200.registry import build_attention
    res = []
    # 2 x 13 x 14 = 364 rows: an even split; 1 x 13 x 15 = 195 rows: uneven shards (97 + 98)
    for bev, bs in (((13, 14), 2), ((13, 15), 1)):
        m = build_attention(lc.CFG_FUSED)
        m.load_state_dict(lc.seeded_state(m, 23))
        m.cuda()
        c = lc.case(seed=5, bs=bs, bev=bev, embed_dims=256)
        for group in (None, dist.group.WORLD):
            m.process_group = group
            m.zero_grad(set_to_none=True)
            e = c["embed"].cuda().requires_grad_(True)
            o = m(e)
            o.backward(c["grad"].cuda())
            res.append([o.detach().cpu(), e.grad.cpu()] + [p.grad.cpu() for _, p in sorted(m.named_parameters())])
        # row-sharded I/O: this rank's rows in, this rank's rows out (no 41 MB all-gather); gathered here for the check
        from vidar_b200.sharding import shard_range
        R = bs * bev[0] * bev[1]
        lo, hi = shard_range(R, rank, world)
        m.process_group = dist.group.WORLD
        m.zero_grad(set_to_none=True)
        e = c["embed"].cuda().view(R, -1)[lo:hi].clone().requires_grad_(True)
        o = m.forward_rows(e, bs, bev[0], bev[1])
        o.backward(c["grad"].cuda().view(R, -1)[lo:hi])
        parts = []
        for t in (o.detach(), e.grad):
            buf = [torch.empty((shard_range(R, r, world)[1] - shard_range(R, r, world)[0], t.shape[1]), device="cuda")
                   for r in range(world)]
            dist.all_gather(buf, t.contiguous())
            parts.append(torch.cat(buf, 0).view(c["embed"].shape).cpu())
        res.append(parts + [p.grad.cpu() for _, p in sorted(m.named_parameters())])
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_fused_module_equals_single_gpu(tmp_path):
    """The whole LatentRendering module (fused projections + core) with rows/cells split over two
    GPUs returns the same replicated output, grad embed and parameter gradients as one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "m.pt")
    mp.spawn(_module_worker, args=(2, 29440 + os.getpid() % 200, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    # per case: [single, sharded (replicated in/out), sharded (row-sharded in/out)]
    for single, sharded in ((res[0], res[1]), (res[0], res[2]), (res[3], res[4]), (res[3], res[5])):
        for i, (a, b) in enumerate(zip(sharded, single)):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-5 * float(b.abs().max()), msg=lambda m: f"tensor {i}: {m}")


def _sca_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tests import module_cases as mc
    import vidar_b200.modules  # noqa: F401
    from vidar_b200 import sharding
    from vidar_b200.registry import build_attention
    res = []
    for bs in (1, 2):
        m = build_attention(mc.SCA_CFG)
        m.load_state_dict(mc.seeded_state(m, 10))
        m.eval().cuda()
        for group in (None, dist.group.WORLD):
            m.set_process_group(group)
            m.zero_grad(set_to_none=True)
            o, gq, gkv = mc.run_module(m, "sca", mc.sca_case(bs=bs), device="cuda")
            if group is not None:
                # camera-sharded: grad wrt the image features is non-zero only for this rank's cameras
                dist.all_reduce(gkv, group=group)
                sharding.allreduce_partial_grads(m, group)
            res.append([o.cpu(), gq.cpu(), gkv.cpu()] + [p.grad.cpu() for _, p in sorted(m.named_parameters())])
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_camera_sharded_sca_equals_single_gpu(tmp_path, world):
    """SpatialCrossAttention with its cameras sharded over `world` GPUs (reduce-scatter of the partial BEV
    slots, output_proj on the rank's rows, ONE all-gather of the BEV grid) returns the single-GPU output,
    input gradients and -- after `allreduce_partial_grads` -- parameter gradients."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "sca.pt")
    mp.spawn(_sca_worker, args=(world, 29240 + os.getpid() % 200 + world, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    for single, sharded in ((res[0], res[1]), (res[2], res[3])):
        for i, (a, b) in enumerate(zip(sharded, single)):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-5 * float(b.abs().max()) + 1e-7, msg=lambda m_: f"tensor {i}: {m_}")
