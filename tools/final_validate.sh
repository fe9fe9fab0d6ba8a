#!/bin/bash
# Round-end validation on the GPU box: parity tests, smoke, both bench arms, ncu launch list + full captures of the top kernels.
# usage: bash tools/final_validate.sh TAG      (results under gpurun_out/)
TAG=${1:-r1}
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "^==|passed|failed|FAILED|rror"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --dump-ops gpurun_out/${TAG}_ops_profile.json > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_n1.json | cut -c1-600
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_reference_arm.json | cut -c1-400
# launch list (shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu_launch.err
python - <<P
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv", errors="replace")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    a = agg.setdefault(r[ki][:64], [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/${TAG}_launch_list_ncu_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
    f.write("(B=32 ViT-L 518px T=1369; the window skips the weight-loading kernels and covers ~2.5 steady-state steps)\n")
    f.write("per-launch times are cold-cache and serialised under the profiler: compare SHARES with the live CUDA-event profile, not absolutes\n\n")
    f.write(f"{'kernel':64s} {'launches':>8s} {'total us':>12s} {'share':>7s}\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:64s} {n:8d} {us:12.1f} {us / tot * 100:6.1f}%\n")
print(open("gpurun_out/${TAG}_launch_list_ncu_summary.txt").read()[:1500])
P
[ -n "$NOFULL" ] && exit 0
# full captures of the top kernels (one launch each)
WHAT=attn timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_attn python tools/prof_conv.py > /dev/null 2>&1
WHAT=conv64skip timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv64_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_conv64 python tools/prof_dec.py > /dev/null 2>&1
ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma2_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_gemm python tools/prof_gemm.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
