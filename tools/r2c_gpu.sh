mkdir -p gpurun_out
T="tests/test_msda_gpu.py tests/test_parity_fullsize_gpu.py tests/test_modules_gpu.py"
VIDAR_MSDA_SLAB=1 timeout 900 python -m pytest $T -q -x 2>&1 | tail -15 > gpurun_out/r2c_pytest_slab1.log
VIDAR_MSDA_SLAB=2 timeout 900 python -m pytest $T -q -x 2>&1 | tail -15 > gpurun_out/r2c_pytest_slab2.log
VIDAR_MSDA_SLAB=2 VIDAR_MSDA_SLAB_FWD=1 timeout 900 python -m pytest $T -q -x 2>&1 | tail -15 > gpurun_out/r2c_pytest_slab2f.log
for m in 0 1 2; do VIDAR_MSDA_SLAB=$m timeout 200 python tools/exp_msda.py >> gpurun_out/r2c_exp.jsonl 2>>gpurun_out/r2c_exp.err; done
VIDAR_MSDA_SLAB=2 VIDAR_MSDA_SLAB_FWD=1 timeout 200 python tools/exp_msda.py >> gpurun_out/r2c_exp.jsonl 2>>gpurun_out/r2c_exp.err
VIDAR_MSDA_SLAB=1 VIDAR_MSDA_SLAB_FWD=1 timeout 200 python tools/exp_msda.py >> gpurun_out/r2c_exp.jsonl 2>>gpurun_out/r2c_exp.err
timeout 200 python tools/debug_dvr_ties.py > gpurun_out/r2c_dvr_ties.jsonl 2>gpurun_out/r2c_dvr.err
VIDAR_DVR_TIE_EPS=1e-6 timeout 200 python tools/debug_dvr_ties.py >> gpurun_out/r2c_dvr_ties.jsonl 2>>gpurun_out/r2c_dvr.err
timeout 600 python -m pytest tests/test_pretrain_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2c_pytest_pretrain.log
timeout 300 python tools/bench_ref_msda.py > gpurun_out/r2c_ref_msda.json 2> gpurun_out/r2c_ref_msda.err
timeout 900 python bench.py --workload pretrain --steps 2 --warmup 1 > gpurun_out/r2c_pretrain.json 2> gpurun_out/r2c_pretrain.err
tail -3 gpurun_out/r2c_pytest_slab1.log gpurun_out/r2c_pytest_slab2.log gpurun_out/r2c_pytest_slab2f.log gpurun_out/r2c_pytest_pretrain.log
cat gpurun_out/r2c_exp.jsonl; cat gpurun_out/r2c_dvr_ties.jsonl | cut -c1-1500; tail -5 gpurun_out/r2c_pretrain.err; cat gpurun_out/r2c_pretrain.json | cut -c1-1500
