mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_multigpu_gpu.py -q -k "sca" 2>&1 | tail -6 > gpurun_out/r2h_pytest_mgpu.log
timeout 400 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err
timeout 400 $TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err
timeout 400 $TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 5 --no-graph > gpurun_out/r2h_bench_n8_nograph.json 2> gpurun_out/r2h_bench_n8_nograph.err
timeout 600 $TR --nproc-per-node 8 --master-port 29604 tools/sweep_multi.py --gpus 8 > gpurun_out/r2h_sweep_n8.json 2> gpurun_out/r2h_sweep_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29605 bench.py --gpus 8 --workload pretrain --steps 2 --warmup 1 > gpurun_out/r2h_pretrain_n8.json 2> gpurun_out/r2h_pretrain_n8.err
tail -4 gpurun_out/r2h_pytest_mgpu.log
python - <<'PY'
import json
for f in ("r2h_bench_n8","r2h_bench_n4","r2h_bench_n8_nograph","r2h_pretrain_n8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read()); print(f, d["ms_per_step"], d.get("launch_mode"), d.get("breakdown_ms"), (d.get("sharded_check") or {}).get("ok"), (d.get("sharded_check") or {}).get("max_over_ranks"), d.get("stage_ms"), (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
try:
    d=json.loads(open("gpurun_out/r2h_sweep_n8.json").read()); print("sweep configs", len(d["configs"])); print([ (c["bev"],c["rays"],c["futures"],round(c["ms_per_step"],3)) for c in d["configs"]])
except Exception as e: print("sweep ERR", e)
PY
tail -3 gpurun_out/r2h_bench_n8.err gpurun_out/r2h_sweep_n8.err gpurun_out/r2h_pretrain_n8.err
