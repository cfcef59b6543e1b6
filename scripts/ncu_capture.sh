#!/usr/bin/env bash
# The ncu recipe behind profiles/ (run on the B200 box, e.g. through gpurun; summaries land in gpurun_out/).
#   scripts/ncu_capture.sh [workload] [tag]      workload in bench.WORKLOADS, default bench_q10k_border
set -euo pipefail
W=${1:-bench_q10k_border}
TAG=${2:-r1}
mkdir -p gpurun_out /tmp/msda_prof
# 1. every launch of a short bench run with its device time (cold cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_${W}.csv \
    python bench.py --steps 5 --warmup 3 --quick > gpurun_out/${TAG}_launches_${W}.log 2>&1
# 2. the two MSDA kernels with the full metric set and source correlation (compiled with -lineinfo)
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 2 -f -o /tmp/msda_prof/prof_${W} \
    python scripts/profile_step.py --workload ${W} --steps 3 > gpurun_out/${TAG}_prof_${W}.log 2>&1
# 3. text exports (the raw report is ~40 MB; gpurun brings back at most 64 MB, so only the exports travel)
ncu -i /tmp/msda_prof/prof_${W}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_${W}.csv
ncu -i /tmp/msda_prof/prof_${W}.ncu-rep --page source --csv --kernel-name regex:msda_bwd > gpurun_out/${TAG}_src_bwd_${W}.csv || true
ncu -i /tmp/msda_prof/prof_${W}.ncu-rep --page source --csv --kernel-name regex:msda_fwd > gpurun_out/${TAG}_src_fwd_${W}.csv || true
gzip -f gpurun_out/${TAG}_src_bwd_${W}.csv gpurun_out/${TAG}_src_fwd_${W}.csv || true
# 4. read them here (no GPU needed):
#    python scripts/summarize_launches.py gpurun_out/${TAG}_launches_${W}.csv
#    python scripts/ncu_source_summary.py <(zcat gpurun_out/${TAG}_src_bwd_${W}.csv.gz) 30
