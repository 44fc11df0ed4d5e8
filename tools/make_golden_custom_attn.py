"""Golden vectors for the detection-decoder attention (decoder.py:132-345), made by running the
REFERENCE class (imported from /root/reference through tools/ref_shim.py) on CPU:

    python tools/make_golden_custom_attn.py        # writes tests/golden/modules_custom.npz

Two cases: 2-component reference points and 4-component reference boxes.
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
    ref_shim.load("modules.decoder")
    reg = pkg.registries["ATTENTION"]
    rec = {}
    for kind, boxes, seed in (("custom", False, 13), ("custom_boxes", True, 14)):
        m = reg.build(mc.CUSTOM_CFG)
        m.load_state_dict(mc.seeded_state(m, seed))
        m.eval()
        out, gq, gkv = mc.run_module(m, kind, mc.custom_case(boxes=boxes))
        rec[f"{kind}_out"], rec[f"{kind}_gq"], rec[f"{kind}_gkv"] = out.numpy(), gq.numpy(), gkv.numpy()
        rec[f"{kind}_params"] = np.array(sorted(m.state_dict().keys()))
        print(kind, tuple(out.shape), float(out.abs().mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "modules_custom.npz"), **rec)


if __name__ == "__main__":
    main()
