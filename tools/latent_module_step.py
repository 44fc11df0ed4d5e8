"""A few LatentRendering module steps at the bench shape (for `ncu` launch lists):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lr_launches.csv \
        python tools/latent_module_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vidar_b200.modules  # noqa: F401,E402
from vidar_b200.registry import build_attention  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    mod = build_attention(bench.LR_CFG).to(dev)
    emb = torch.randn(1, 200, 200, 256, device=dev)
    gemb = torch.randn(1, 200, 200, 256, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(int(os.environ.get("STEPS", "3"))):
        flush.zero_()                       # evict the 126 MB L2 like the MSDA stage does in the bench
        e = emb.detach().requires_grad_(True)
        mod.zero_grad(set_to_none=True)
        mod(e).backward(gemb)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
