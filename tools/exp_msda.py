"""Experiment driver: MSDA forward / backward time at the cfg2 shapes (6 cameras, dense 40000 and
rebatched 10240 queries) for the library as built, one JSON line.  Knobs are environment variables read
by the library (VIDAR_MSDA_STREAM_BYTES) so variants are separate processes:
    for v in 0xffffffff 262144 1000000 100000; do VIDAR_MSDA_STREAM_BYTES=$v python tools/exp_msda.py; done"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidar_b200 import msda, synthetic  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {"stream_bytes": os.environ.get("VIDAR_MSDA_STREAM_BYTES", "default"), "slab": os.environ.get("VIDAR_MSDA_SLAB", "default"),
       "slab_fwd": os.environ.get("VIDAR_MSDA_SLAB_FWD", "0")}
for tag, rows in (("dense", None), ("rebatched", 10240)):
    d = synthetic.sca_like_inputs(dev, rows=rows)
    cams = d["value"].shape[0]
    gv = [torch.zeros_like(d["value"][c:c + 1]) for c in range(cams)]
    gl, ga = torch.empty_like(d["loc"][:1]), torch.empty_like(d["attn"][:1])
    per = [dict(value=d["value"][c:c + 1], loc=d["loc"][c:c + 1].contiguous(), attn=d["attn"][c:c + 1].contiguous(),
                go=d["grad_out"][c:c + 1].contiguous()) for c in range(cams)]

    def fwd():
        for s in per:
            msda.ext_module.ms_deform_attn_forward(s["value"], d["shapes"], d["lsi"], s["loc"], s["attn"], im2col_step=64)

    def bwd():
        for c, s in enumerate(per):
            msda.ext_module.ms_deform_attn_backward(s["value"], d["shapes"], d["lsi"], s["loc"], s["attn"], s["go"],
                                                    gv[c], gl, ga, im2col_step=64)
    res[tag] = {"fwd_ms": timeit(fwd), "bwd_ms": timeit(bwd)}
    del d, per, gv, gl, ga
    torch.cuda.empty_cache()
print(json.dumps(res))
