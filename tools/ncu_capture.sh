#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun; one GPU).  Usage: bash tools/ncu_capture.sh TAG
#   1. launch list of a short bench run (kernel SHARES of one MPC tick; per-launch times are cold-cache and serialised)
#   2. one `--set full` capture each of the Riccati kernel and of the two evaluation kernels (4th launch of each: warm state)
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
COMMON="--clock-control none"
ncu $COMMON --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_batch592.csv \
    python bench.py --steps 2 --warmup 1 --batch 592 --prep-iters 6 --no-cpu-baseline --latency-ticks 0 > $OUT/${TAG}_launches_bench.log 2>&1
ncu $COMMON --set full --import-source on --kernel-name-base demangled -k regex:k_riccati --launch-skip 3 -c 1 -f -o $OUT/${TAG}_prof_riccati \
    python tools/profile_driver.py 592 3 > $OUT/${TAG}_ncu_riccati.log 2>&1
ncu $COMMON --set full --import-source on --kernel-name-base demangled -k "regex:k_eval<.*(true|1)>" --launch-skip 3 -c 1 -f -o $OUT/${TAG}_prof_evald \
    python tools/profile_driver.py 592 3 > $OUT/${TAG}_ncu_evald.log 2>&1
ncu $COMMON --set full --import-source on --kernel-name-base demangled -k "regex:k_eval<.*(false|0)>" --launch-skip 3 -c 1 -f -o $OUT/${TAG}_prof_evalv \
    python tools/profile_driver.py 592 3 > $OUT/${TAG}_ncu_evalv.log 2>&1
ls -la $OUT/${TAG}_*
tail -2 $OUT/${TAG}_ncu_riccati.log $OUT/${TAG}_ncu_evald.log $OUT/${TAG}_ncu_evalv.log
