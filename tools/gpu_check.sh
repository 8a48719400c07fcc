#!/bin/bash
# GPU-box check used during kernel work: parity tests, in-kernel phase counters, a short bench.  Usage: bash tools/gpu_check.sh TAG
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gputests.log 2>&1; tail -3 $OUT/${TAG}_gputests.log
MPCB200_LIB=mpc_benchmark_b200/libmpcb200_phase.so python tools/phase_timing.py > $OUT/${TAG}_phase.log 2>&1; tail -6 $OUT/${TAG}_phase.log
python bench.py --no-cpu-baseline --latency-ticks 100 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 400 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernel_ms_per_step"], d.get("latency",{}).get("p50_ms"), d["roofline"]["frac"])
PY
