"""BASELINE.json configs[4]: the sweep BEV {100,200,400}^2 x rays {10k,30k,100k} x futures {1,3,6} through the
product path at N GPUs (torchrun), one JSON document on rank 0's stdout:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        tools/sweep_multi.py --gpus 8 > gpurun_out/sweep_n8.json
Per configuration: ms/step (CUDA graph replay, max over ranks), rays/s, the per-stage breakdown (eager pass) and
the dominant kernel's algorithmic GB/s against the measured HBM peak (SURVEY.md 8d byte counts)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="the diagonal only (3 configurations)")
    a = ap.parse_args()
    a.graph = True
    rank = int(os.environ.get("RANK", "0"))
    peak, _ = bench.peaks()
    grid = [(b, r, f) for b in (100, 200, 400) for r in (10000, 30000, 100000) for f in (1, 3, 6)]
    if a.quick:
        grid = [(100, 10000, 1), (200, 30000, 3), (400, 100000, 6)]
    out = {"n_gpus": a.gpus, "hbm_peak_GBps": peak, "configs": []}
    for bev, rays, frames in grid:
        bench.set_workload(bev, rays, frames)
        r = bench.run_ours(a, light=True)
        torch.cuda.empty_cache()
        if rank == 0:
            rows = r.pop("rows_this_rank")
            cams = max(1, min(bench.NUM_CAMS, -(-bench.NUM_CAMS // a.gpus) + 1))
            fwd_b, bwd_b = bench.msda_algorithmic_bytes_q(rows, cams, bev * bev)
            r.update(bev=bev, rays=rays, futures=frames,
                     msda_bwd_GBps=bwd_b / (r["breakdown_ms"]["msda_bwd"] * 1e-3) / 1e9,
                     msda_fwd_GBps=fwd_b / (r["breakdown_ms"]["msda_fwd"] * 1e-3) / 1e9)
            r["msda_bwd_pct_of_hbm_peak"] = 100 * r["msda_bwd_GBps"] / peak
            out["configs"].append(r)
    if rank == 0:
        os.write(bench._REAL_STDOUT, (json.dumps(out, indent=1) + "\n").encode())
    bench._finish(a.gpus)      # multi-rank: leave without process-group teardown (see bench._finish)


if __name__ == "__main__":
    main()
