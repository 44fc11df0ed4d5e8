"""Golden vectors for the NN / Chamfer op from the REFERENCE's own code run on CPU here:
oracle/_ref/ref_knn_cpu.so (chamferdist ext.cpp + knn_cpu.cpp, unmodified) exposed as
`chamferdist._C`, with the reference chamfer.py (ChamferDistance, knn_points) on top.
Writes tests/golden/knn.npz."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests.knn_cases import case  # noqa: E402

REF = "/root/reference/third_lib/chamfer_dist/chamferdist/chamferdist"


def main():
    build_ref.build_knn_cpu()
    C = build_ref.load("ref_knn_cpu")
    pkg = types.ModuleType("chamferdist")
    pkg._C = C
    pkg.__path__ = []
    sys.modules["chamferdist"] = pkg
    spec = importlib.util.spec_from_file_location("chamferdist.chamfer", os.path.join(REF, "chamfer.py"))
    ch = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ch)
    c = case()
    a = c["a"].clone().requires_grad_(True)
    b = c["b"].clone().requires_grad_(True)
    k = ch.knn_points(a, b, lengths1=c["la"], lengths2=c["lb"], K=1)
    (k.dists * c["g"]).sum().backward()
    rec = dict(dists=k.dists.detach().numpy(), idx=k.idx.numpy(), ga=a.grad.numpy(), gb=b.grad.numpy())
    cd = ch.ChamferDistance()
    f, bwd, info = cd(c["a"], c["b"], bidirectional=True, reduction="sum")
    rec.update(cham_fwd=f.numpy(), cham_bwd=bwd.numpy(), info_fd=info[0].numpy(), info_fi=info[1].numpy(),
               info_bd=info[2].numpy(), info_bi=info[3].numpy())
    m, _ = cd(c["a"], c["b"], reduction="mean")
    r, _ = cd(c["a"], c["b"], reverse=True, reduction=None)
    rec.update(cham_mean=m.numpy(), cham_rev=r.numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "knn.npz"), **rec)
    print("ok", float(f), float(bwd))


if __name__ == "__main__":
    main()
