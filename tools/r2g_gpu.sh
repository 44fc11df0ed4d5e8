mkdir -p gpurun_out
for v in 0xffffffff 100000 262144; do VIDAR_MSDA_STREAM_BYTES=$v timeout 200 python tools/exp_msda.py >> gpurun_out/r2g_exp.jsonl 2>>gpurun_out/r2g_exp.err; done
for v in 0xffffffff 100000; do VIDAR_MSDA_STREAM_BYTES=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-graph 2>>gpurun_out/r2g_exp.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'stream': '$v', 'ms': d['ms_per_step'], 'parts': d['breakdown_ms']}))" >> gpurun_out/r2g_bench.jsonl; done
cat gpurun_out/r2g_exp.jsonl gpurun_out/r2g_bench.jsonl
