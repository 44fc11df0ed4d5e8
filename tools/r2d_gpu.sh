mkdir -p gpurun_out
timeout 300 python tools/debug_dvr_ties.py > gpurun_out/r2d_dvr_ties.json 2>gpurun_out/r2d_dvr.err
timeout 600 python tools/make_golden_dvr.py ties > gpurun_out/r2d_golden.log 2>&1
timeout 600 python -m pytest tests/test_msda_gpu.py tests/test_dvr_gpu.py tests/test_latent_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-graph > gpurun_out/r2d_bench_nograph.json 2>> gpurun_out/r2d_bench.err
cat gpurun_out/r2d_dvr_ties.json; tail -3 gpurun_out/r2d_golden.log; tail -4 gpurun_out/r2d_pytest.log; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2d_bench.json","gpurun_out/r2d_bench_nograph.json"):
    try:
        d=json.loads(open(f).read()); print(f, d["ms_per_step"], d["launch_mode"], d["breakdown_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
