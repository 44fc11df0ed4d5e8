"""bench.py's reference arm (the CPU leg the driver runs beside the GPU arm) on a GPU-less host: one JSON line on
stdout, the contract's keys, and the SAME `config` object as the GPU arm would print for that GPU count."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gpus,rank", [(1, 0), (8, 0), (8, 3)])
def test_reference_arm_line(gpus, rank):
    env = dict(os.environ, VIDAR_REF_BUDGET_S="2", VIDAR_REF_FULL_STEP="0", RANK=str(rank), WORLD_SIZE=str(gpus))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus),
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    if rank != 0:                      # under torchrun only rank 0 works and prints
        assert p.stdout.strip() == ""
        return
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == gpus and line["steps"] == 2 and line["warmup"] == 1
    assert line["unit"] == "rays/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["estimated"] is True and 0 < line["sample_fraction"] < 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config(gpus) and line["metric"] == bench.METRIC


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert p.returncode != 0 and p.stdout.strip() == "" and "no CPU fallback" in p.stderr


def test_multi_rank_exit_leaves_after_the_line_without_teardown(tmp_path):
    """bench._finish: every rank passes one barrier after rank 0's line and leaves with status 0 without running
    process-group destructors (2 gloo ranks here; the CUDA synchronize is a no-op on this host)."""
    code = (
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "torch.cuda.synchronize = lambda *a, **k: None\n"
        "dist.init_process_group('gloo')\n"
        "if dist.get_rank() == 0:\n"
        "    bench._emit('{\"ok\": true}')\n"
        "bench._finish(dist.get_world_size())\n"
        "print('not reached')\n")
    port = 29900 + os.getpid() % 90
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT))
    outs = [p.communicate(timeout=300) for p in procs]
    assert [p.returncode for p in procs] == [0, 0], outs
    assert outs[0][0].strip() == '{"ok": true}' and outs[1][0].strip() == ""
