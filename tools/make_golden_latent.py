"""Golden vectors for LatentRendering from the REFERENCE class run on CPU in this container
(latent_rendering.py:37-162 through tools/ref_shim.py).  Writes tests/golden/latent_rendering.npz.
The reference builds its grids with device='cuda' by default (get_bev_grids); it is called
here with torch's default device patched to CPU via the function's own `device` default."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import latent_cases as lc  # noqa: E402
from tools import ref_shim  # noqa: E402


def main():
    mod = ref_shim.load("modules.ray_operations.latent_rendering")
    # get_bev_grids(H, W, bs, device='cuda', ...) hard-codes CUDA: rebind the default to CPU
    f = mod.get_bev_grids
    mod.get_bev_grids = lambda H, W, bs=1, device="cpu", dtype=torch.float, offset=0.5: f(H, W, bs, "cpu", dtype, offset)
    rec = {}
    for tag, cfg, seed in (("sig", lc.CFG, 20), ("exp", lc.CFG_EXP, 21), ("d1", lc.CFG_D1, 22)):
        kw = dict(cfg)
        kw.pop("type")
        m = mod.LatentRendering(**kw)
        m.load_state_dict(lc.seeded_state(m, seed))
        c = lc.case()
        e = c["embed"].clone().requires_grad_(True)
        out = m(e)
        out.backward(c["grad"])
        rec[f"{tag}_out"], rec[f"{tag}_gembed"] = out.detach().numpy(), e.grad.numpy()
        rec[f"{tag}_params"] = np.array(sorted(m.state_dict().keys()))
        for n, p in m.named_parameters():
            if n.startswith("lora_b.weight") or n.endswith("head.0.weight") or n.endswith("head.3.weight"):
                rec[f"{tag}_g_{n}"] = p.grad.numpy()
        print(tag, tuple(out.shape), float(out.abs().mean()))
    if "--fused-only" not in sys.argv:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "latent_rendering.npz"), **rec)
    # shapes of the fused projection path: every parameter gradient is stored
    rec = {}
    for tag, cfg, seed in (("fused", lc.CFG_FUSED, 23), ("fused_exp", lc.CFG_FUSED_EXP, 24)):
        kw = dict(cfg)
        kw.pop("type")
        m = mod.LatentRendering(**kw)
        m.load_state_dict(lc.seeded_state(m, seed))
        c = lc.case(seed=5, bev=lc.BEV_FUSED, embed_dims=cfg["embed_dims"])
        e = c["embed"].clone().requires_grad_(True)
        out = m(e)
        out.backward(c["grad"])
        rec[f"{tag}_out"], rec[f"{tag}_gembed"] = out.detach().numpy(), e.grad.numpy()
        rec[f"{tag}_params"] = np.array(sorted(m.state_dict().keys()))
        for n, p in m.named_parameters():
            rec[f"{tag}_g_{n}"] = p.grad.numpy()
        print(tag, tuple(out.shape), float(out.abs().mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "latent_fused.npz"), **rec)


if __name__ == "__main__":
    main()
