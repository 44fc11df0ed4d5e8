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
