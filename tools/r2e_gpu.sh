mkdir -p gpurun_out
timeout 300 python tools/debug_dvr_ties.py > gpurun_out/r2e_dvr_ties.json 2>gpurun_out/r2e_dvr.err
timeout 900 python -m pytest tests/test_dvr_gpu.py tests/test_multigpu_gpu.py tests/test_pretrain_gpu.py tests/test_head_gpu.py -q 2>&1 | tail -25 > gpurun_out/r2e_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload pretrain --steps 2 --warmup 1 > gpurun_out/r2e_pretrain_n2.json 2> gpurun_out/r2e_pretrain_n2.err
cat gpurun_out/r2e_dvr_ties.json | cut -c1-2500; tail -8 gpurun_out/r2e_pytest.log; tail -3 gpurun_out/r2e_bench_n2.err; tail -3 gpurun_out/r2e_pretrain_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2e_bench_n2.json","gpurun_out/r2e_pretrain_n2.json"):
    try:
        d=json.loads(open(f).read()); print(f, d["ms_per_step"], d.get("launch_mode"), d.get("breakdown_ms"), d.get("sharded_check"), d.get("stage_ms"))
    except Exception as e: print(f, "ERR", e)
PY
