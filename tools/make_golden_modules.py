"""Golden vectors for the attention MODULES, made by running the REFERENCE's own classes
(imported from /root/reference through tools/ref_shim.py) on CPU in this container:

    python tools/make_golden_modules.py        # writes tests/golden/modules.npz

Inputs and weights are regenerated from seeds by tests/module_cases.py, so the file holds
outputs and gradients only.  On CPU the reference classes take their own
`multi_scale_deformable_attn_pytorch` branch (spatial_cross_attention.py:392-394 etc.).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import module_cases as mc  # noqa: E402
from tools import ref_shim  # noqa: E402


def main():
    torch.manual_seed(0)
    pkg = ref_shim.install()
    ref_shim.load("modules.spatial_cross_attention")
    ref_shim.load("modules.temporal_self_attention")
    ref_shim.load("modules.vidar_decoder")
    reg = pkg.registries["ATTENTION"]
    rec = {}
    for kind, cfg, case, seed in (("sca", mc.SCA_CFG, mc.sca_case(), 10), ("tsa", mc.TSA_CFG, mc.tsa_case(), 11),
                                  ("pred", mc.PRED_CFG, mc.pred_case(), 12)):
        m = reg.build(cfg)
        m.load_state_dict(mc.seeded_state(m, seed))
        m.eval()
        out, gq, gkv = mc.run_module(m, kind, case)
        rec[f"{kind}_out"], rec[f"{kind}_gq"] = out.numpy(), gq.numpy()
        rec[f"{kind}_gkv_s6"] = gkv[:, ::6].numpy()      # every 6th key row keeps the file small
        rec[f"{kind}_params"] = np.array(sorted(m.state_dict().keys()))
        print(kind, tuple(out.shape), float(out.abs().mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "modules.npz"), **rec)


if __name__ == "__main__":
    main()
