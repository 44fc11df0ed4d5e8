"""Row-sharded encoder layers (vidar_b200/pretrain.py::EncoderLayer.forward_rows, SURVEY.md 8e: TSA / norms / FFN on
the rank's BEV rows with the value replicated, camera-sharded SpatialCrossAttention returning rows) over gloo, world 2,
against the plain single-process layers: output, gradients of every replicated input, gradients of the image
features and -- after `allreduce_partial_grads` -- every parameter gradient.

Host logic only: the CUDA sampling op is replaced by the CPU oracle (`msda_apply`), and the camera-sharded slot
kernel by the reference data flow restricted to the rank's cameras."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

BEV = (8, 8)
CAM_LEVELS = ((4, 6), (2, 3), (1, 2), (1, 1))
CAMS, D = 6, 4


def _inputs():
    from vidar_b200.pretrain import EMBED
    g = torch.Generator().manual_seed(7)
    Q = BEV[0] * BEV[1]
    shapes = torch.tensor(CAM_LEVELS, dtype=torch.int64)
    hw = shapes[:, 0] * shapes[:, 1]
    lsi = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
    K = int(hw.sum())
    mask = torch.rand(CAMS, 1, Q, D, generator=g) < 0.35
    mask[:, :, 5] = False                                   # a pillar no camera sees
    return dict(
        x=torch.randn(1, Q, EMBED, generator=g), pos=0.1 * torch.randn(1, Q, EMBED, generator=g),
        prev=torch.randn(1, Q, EMBED, generator=g), feats=torch.randn(CAMS, K, 1, EMBED, generator=g),
        ref_cam=torch.rand(CAMS, 1, Q, D, 2, generator=g) * 1.2 - 0.1, mask=mask,
        ref_2d=torch.rand(2, Q, 1, 2, generator=g), shapes=shapes, lsi=lsi,
        bev_shapes=torch.tensor([BEV], dtype=torch.int64), bev_lsi=torch.tensor([0]),
        gout=torch.randn(1, Q, EMBED, generator=g))


def _layers():
    from vidar_b200.pretrain import EncoderLayer
    torch.manual_seed(11)
    layers = torch.nn.ModuleList([EncoderLayer(False, BEV) for _ in range(2)])
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():                                   # the default init zeroes the offset / weight Linears
        for p in layers.parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5)
    return layers.eval()                                    # dropout off


def _patch_oracle():
    from oracle import msda_ref
    from vidar_b200.modules import deform_attn

    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_ref.msda_grid_sample(value, shapes, loc, attn)
    deform_attn.msda_apply = apply


def _run(layers, d, group):
    from vidar_b200 import sharding
    x0, pos, feats = (d[k].clone().requires_grad_(True) for k in ("x", "pos", "feats"))
    x, p = x0, pos
    if group is not None:
        x, p = sharding.sum_grad(x, group), sharding.sum_grad(p, group)
    Q = x.shape[1]
    for i, layer in enumerate(layers):
        pv = torch.stack([d["prev"], x], 1).reshape(2, Q, -1)
        args = (x, feats, p, d["ref_2d"], pv, d["shapes"], d["lsi"], d["ref_cam"], d["mask"], d["bev_shapes"], d["bev_lsi"])
        x = layer(*args) if group is None else layer.forward_rows(*args, group, last=i == len(layers) - 1)
    (x * d["gout"]).sum().backward()
    return x.detach(), x0.grad, pos.grad, feats.grad


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vidar_b200 import sca, sharding
    from vidar_b200.modules.deform_attn import SpatialCrossAttention
    _patch_oracle()

    def slots_of_my_cameras(self, query, value, ref_cam, bev_mask, shapes, lsi):
        w, r = self._world()
        mine = sca.plan_cameras(sca.unit_plan(w, r, value.shape[0]))
        m = torch.zeros_like(bev_mask)
        m[mine] = bev_mask[mine]
        part = self._slots_rebatch(query, value, value, ref_cam, m, shapes, lsi)       # sum over my cameras / #mine
        cnt = lambda mm: torch.clamp((mm.sum(-1) > 0).permute(1, 2, 0).sum(-1), min=1.0)
        return part * (cnt(m) / cnt(bev_mask))[..., None]                              # ... / #all cameras
    SpatialCrossAttention._fusable = lambda self, *a: True
    SpatialCrossAttention._slots_fused = slots_of_my_cameras

    layers, d = _layers(), _inputs()
    group = dist.group.WORLD
    for layer in layers:
        layer.cross_attn.set_process_group(group)
        for m in (layer.self_attn, layer.norms, layer.ffn):
            sharding.mark_partial(m)
    y, gx, gpos, gfeats = _run(layers, d, group)
    dist.all_reduce(gfeats)                                 # each rank holds the gradient of its cameras' features
    sharding.allreduce_partial_grads(layers, group)
    torch.save(dict(y=y, gx=gx, gpos=gpos, gfeats=gfeats, params={n: p.grad for n, p in layers.named_parameters()}),
               f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_encoder_layers_equal_plain_layers(tmp_path):
    world = 2
    out = str(tmp_path / "enc.pt")
    mp.spawn(_worker, args=(world, 29100 + os.getpid() % 300, out), nprocs=world, join=True)
    _patch_oracle()
    layers, d = _layers(), _inputs()
    y, gx, gpos, gfeats = _run(layers, d, None)
    want = dict(y=y, gx=gx, gpos=gpos, gfeats=gfeats)
    for rank in range(world):                               # every rank ends with the full, identical results
        got = torch.load(f"{out}.{rank}", weights_only=False)
        for k, w in want.items():
            torch.testing.assert_close(got[k], w, rtol=1e-4, atol=1e-5 * float(w.abs().max()), msg=lambda m: f"rank {rank} {k}: {m}")
        for n, p in layers.named_parameters():
            g = got["params"][n]
            assert g is not None, n
            torch.testing.assert_close(g, p.grad, rtol=1e-4, atol=2e-5 * float(p.grad.abs().max()) + 1e-9,
                                       msg=lambda m: f"rank {rank} param {n}: {m}")


# ---- decoder stack: two future frames chained through `run_decoder`, every layer output also read by a replicated head
def _decoder_layers():
    from vidar_b200.pretrain import DecoderLayer
    torch.manual_seed(21)
    layers = torch.nn.ModuleList([DecoderLayer() for _ in range(2)])
    g = torch.Generator().manual_seed(22)
    with torch.no_grad():
        for p in layers.parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5)
    return layers.eval()


def _run_decoder(layers, d, group):
    from vidar_b200 import pretrain
    q0, pos, prev0 = (d[k].clone().requires_grad_(True) for k in ("x", "pos", "prev"))
    head = torch.nn.Linear(q0.shape[-1], 8)
    with torch.no_grad():
        head.weight.copy_(torch.linspace(-1, 1, head.weight.numel()).view_as(head.weight))
        head.bias.zero_()
    ref = d["ref_2d"][:1]
    loss, prev = 0.0, prev0
    for frame in range(2):
        inter, last = pretrain.run_decoder(layers, q0, prev, pos, ref, d["bev_shapes"], d["bev_lsi"], group)
        loss = loss + (head(inter) * (frame + 1)).square().mean()          # replicated consumer of every layer output
        prev = last + 0.5                                                   # next frame's previous BEV
    (loss + (last * d["gout"]).sum()).backward()
    return loss.detach(), q0.grad, pos.grad, prev0.grad, head.weight.grad


def _decoder_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vidar_b200 import sharding
    _patch_oracle()
    layers, d = _decoder_layers(), _inputs()
    sharding.mark_partial(layers)
    res = _run_decoder(layers, d, dist.group.WORLD)
    sharding.allreduce_partial_grads(layers, dist.group.WORLD)
    torch.save(dict(res=res, params={n: p.grad for n, p in layers.named_parameters()}), f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_decoder_stack_equals_plain_stack(tmp_path):
    world = 2
    out = str(tmp_path / "dec.pt")
    mp.spawn(_decoder_worker, args=(world, 29450 + os.getpid() % 300, out), nprocs=world, join=True)
    _patch_oracle()
    layers, d = _decoder_layers(), _inputs()
    want = _run_decoder(layers, d, None)
    names = ("loss", "grad_query", "grad_pos", "grad_prev", "grad_head_weight")
    for rank in range(world):
        got = torch.load(f"{out}.{rank}", weights_only=False)
        for k, a, w in zip(names, got["res"], want):
            torch.testing.assert_close(a, w, rtol=1e-4, atol=1e-5 * float(w.abs().max()) + 1e-9, msg=lambda m: f"rank {rank} {k}: {m}")
        for n, p in layers.named_parameters():
            torch.testing.assert_close(got["params"][n], p.grad, rtol=1e-4, atol=2e-5 * float(p.grad.abs().max()) + 1e-9,
                                       msg=lambda m: f"rank {rank} param {n}: {m}")
