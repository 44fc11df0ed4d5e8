#!/bin/bash
# ncu evidence for one round, run on the GPU box (one GPU):
#     gpurun --timeout 900 -- 'bash tools/profile_round.sh r01'
# then here:  python tools/summarize_profiles.py r01   (-> profiles/r01_*.txt, roofline_traffic.json)
# 1. launch list of a short bench run (per-launch durations: compare SHARES, launches are serialised and cold)
# 2. one `--set full` capture of every vidar_b200 kernel of ONE step (bench.py brackets it with
#    cudaProfilerStart/Stop when VIDAR_BENCH_PROFILE=1)
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
VIDAR_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
VIDAR_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:msda_|latent_|proj_|ray_|render_|point_sampling|nn_|dvxlr|list_scatter" -f -o gpurun_out/${TAG}_step python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_step.log 2>&1
ls -la gpurun_out/${TAG}_step.ncu-rep gpurun_out/${TAG}_launches.csv
