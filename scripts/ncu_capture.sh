#!/usr/bin/env bash
# The ncu recipe behind profiles/ (run on the B200 box, e.g. through gpurun; raw reports land in gpurun_out/).
#   scripts/ncu_capture.sh [workload]        workload in bench.WORKLOADS, default bench_q10k_border
set -euo pipefail
W=${1:-bench_q10k_border}
mkdir -p gpurun_out
# 1. every launch of a short bench run with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${W}.csv \
    python bench.py --steps 5 --warmup 3 --quick > gpurun_out/launches_${W}.log 2>&1
# 2. the two MSDA kernels with the full metric set and source correlation (compiled with -lineinfo)
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 2 -o gpurun_out/prof_${W} \
    python scripts/profile_step.py --workload ${W} --steps 3 > gpurun_out/prof_${W}.log 2>&1
# 3. read them back (works without a GPU):
#    ncu -i gpurun_out/prof_${W}.ncu-rep --page raw --csv > raw.csv
#    ncu -i gpurun_out/prof_${W}.ncu-rep --page source --csv --kernel-name regex:msda_bwd > src.csv
#    python scripts/ncu_source_summary.py src.csv 30 ; python scripts/summarize_launches.py gpurun_out/launches_${W}.csv
