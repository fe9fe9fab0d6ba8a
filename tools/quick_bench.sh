#!/bin/bash
# usage (GPU box): bash tools/quick_bench.sh TAG  -> conv/op tests + bench with per-op profile, short summary on stdout
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["latency"]["p50_ms"])
for k in ("roofline","roofline_decoder","roofline_attention"): print(k, round(d[k]["achieved"],1), round(d[k]["ms_per_step"],3))
P
tail -2 gpurun_out/bench_$TAG.err
