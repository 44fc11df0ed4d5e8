"""Experiment: MSDA fwd/bwd time per pyramid level (same sample count, single level of varying size)."""
import sys, os, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidar_b200 import msda
dev = torch.device("cuda:0")
def run(level, Q=40000, H=8, C=32, P=32):
    g = torch.Generator(device=dev).manual_seed(0)
    h, w = level
    K = h * w
    value = torch.randn(1, K, H, C, device=dev, generator=g)
    n = int(Q ** 0.5)
    iy, ix = torch.meshgrid(torch.arange(n, device=dev), torch.arange(n, device=dev), indexing="ij")
    ref = torch.stack([(ix.reshape(-1) + 0.5) / n, (iy.reshape(-1) + 0.5) / n], -1).view(1, Q, 1, 1, 1, 2)
    loc = (ref + 4.0 * torch.randn(1, Q, H, 1, P, 2, device=dev, generator=g) / torch.tensor([w, h], device=dev)).contiguous()
    attn = torch.softmax(torch.randn(1, Q, H, P, device=dev, generator=g), -1).view(1, Q, H, 1, P).contiguous()
    go = torch.randn(1, Q, H * C, device=dev, generator=g)
    shapes = torch.tensor([[h, w]], device=dev); lsi = torch.tensor([0], device=dev)
    gv = torch.zeros_like(value); gl = torch.empty_like(loc); ga = torch.empty_like(attn)
    def f(): return msda.ext_module.ms_deform_attn_forward(value, shapes, lsi, loc, attn, im2col_step=64)
    def b(): msda.ext_module.ms_deform_attn_backward(value, shapes, lsi, loc, attn, go, gv, gl, ga, im2col_step=64)
    out = {}
    for name, fn in (("fwd", f), ("bwd", b)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize()
        out[name] = a.elapsed_time(e) / 10
    return out
res = {str(l): run(l) for l in ((116, 200), (58, 100), (29, 50), (15, 25))}
print(json.dumps(res, indent=1))
