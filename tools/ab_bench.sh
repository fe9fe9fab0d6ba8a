#!/bin/bash
# A/B of two builds on the SAME box: bash tools/ab_bench.sh tools/bin/libA.so tools/bin/libB.so [rounds]
A=$1; B=$2; R=${3:-2}
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for tag in A B; do
    lib=$A; [ $tag = B ] && lib=$B
    cp $lib moge_b200/_lib/libmoge_b200.so
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_$tag$r.json > gpurun_out/bench_$tag$r.json 2> gpurun_out/bench_$tag$r.err
    python - <<P
import json
d=json.load(open("gpurun_out/bench_$tag$r.json"))
print("$tag$r", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms | p50", round(d["latency"]["p50_ms"],3), "| gemm", round(d["roofline"]["ms_per_step"],2), "dec", round(d["roofline_decoder"]["ms_per_step"],2), "att", round(d["roofline_attention"]["ms_per_step"],2), "| clk", d["clocks"]["sm_mhz"])
P
  done
done
