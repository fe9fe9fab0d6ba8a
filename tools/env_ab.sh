#!/bin/bash
# A/B of an environment switch on the SAME box and build: bash tools/env_ab.sh VAR "v1 v2 ..." [rounds]
VAR=$1; VALS=$2; R=${3:-2}
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for v in $VALS; do
    env $VAR=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_${VAR}_$v.json > gpurun_out/bench_${VAR}_$v.json 2> gpurun_out/bench_${VAR}_$v.err
    python - <<P
import json
d=json.load(open("gpurun_out/bench_${VAR}_$v.json"))
print("$VAR=$v r$r", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms | p50", round(d["latency"]["p50_ms"],3), "| gemm", round(d["roofline"]["ms_per_step"],2), "dec", round(d["roofline_decoder"]["ms_per_step"],2), "att", round(d["roofline_attention"]["ms_per_step"],2), "| clk", d["clocks"]["sm_mhz"])
P
  done
done
