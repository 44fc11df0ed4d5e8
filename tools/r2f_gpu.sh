mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_linear_tc_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2f_pytest_linear.log
tail -6 gpurun_out/r2f_pytest_linear.log
timeout 120 python tools/bench_linear_tc.py > gpurun_out/r2f_linear.json 2> gpurun_out/r2f_linear.err; cat gpurun_out/r2f_linear.json; tail -2 gpurun_out/r2f_linear.err
timeout 300 python -m pytest tests/test_dvr_gpu.py -q 2>&1 | tail -4 > gpurun_out/r2f_pytest_dvr.log; tail -3 gpurun_out/r2f_pytest_dvr.log
timeout 600 python tools/make_golden_dvr.py ties > gpurun_out/r2f_golden.log 2>&1
bash tools/profile_round.sh r02 > gpurun_out/r2f_profile.log 2>&1; tail -3 gpurun_out/r2f_profile.log
