#!/bin/bash
# Round-end validation + evidence on the GPU box (one GPU): parity tests, smoke, both bench arms, configs 3 / 5, attribution,
# ncu launch list + full captures of the dominant kernels taken from the REAL launch list.   usage: bash tools/final_validate.sh TAG
# QUICK=1: parity tests, smoke, the default bench line and full captures of the decoder kernels only (a re-validation after a
# decoder-only change; ~6 minutes).   NONCU=1: no ncu at all.
TAG=${1:-r2}
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "^==|passed|failed|FAILED|rror"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/${TAG}_ops_profile.json --gpu-baseline-kernels gpurun_out/${TAG}_gpu_baseline_kernels.json > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_n1.json
if [ -z "$QUICK" ]; then
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --config 3 --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_config3.err; echo "config3 rc=$?"
timeout 900 python bench.py --config 5 --steps 4 > gpurun_out/${TAG}_bench_config5.json 2> gpurun_out/${TAG}_config5.err; echo "config5 rc=$?"
timeout 600 python bench.py --tokens 3600 --batch 8 --steps 5 --warmup 2 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${TAG}_bench_n1_tokens3600_batch8.json 2>/dev/null; echo "t3600 rc=$?"
timeout 600 python tools/error_attribution.py --out gpurun_out/${TAG}_error_attribution.json > /dev/null 2>&1; echo "attribution rc=$?"
fi
[ -n "$NONCU" ] && exit 0
cap() { timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 python tools/prof_model.py > /dev/null 2>&1; ls -la gpurun_out/${TAG}_$1.ncu-rep 2>&1 | tail -1; }
if [ -n "$QUICK" ]; then
  cap conv64 'conv64_kernel<\(int\)64,' 16          # res_b.neck.l3 of the second forward (two MMA-issuing warps)
  cap convh128 'convh_kernel<\(int\)128,' 15        # res_a.neck.l2 of the second forward
  cap headout 'conv64_kernel<\(int\)16,' 3
  for f in gpurun_out/${TAG}_*.ncu-rep; do python tools/ncu_summary.py $f > ${f%.ncu-rep}.txt 2>/dev/null; done
  ls gpurun_out/${TAG}_*.txt
  exit 0
fi
# launch list (shares): default product path
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu_launch.err
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
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline\n")
    f.write("(default product path: LayerNorm fold, neck fold, persistent attention; B=32 ViT-L 518px T=1369; the window skips the weight-loading kernels)\n")
    f.write("per-launch times are cold-cache and serialised under the profiler: compare SHARES with the live CUDA-event profile, not absolutes\n\n")
    f.write(f"{'kernel':64s} {'launches':>8s} {'total us':>12s} {'share':>7s}\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:64s} {n:8d} {us:12.1f} {us / tot * 100:6.1f}%\n")
print(open("gpurun_out/${TAG}_launch_list_ncu_summary.txt").read()[:1800])
P
# full captures, one launch each, from the real launch list (second forward)
cap gemm_qkv 'umma2_kernel<\(int\)0,' 24          # qkv launch of the second forward (24 per forward)
cap gemm_proj 'umma2_kernel<\(int\)2,' 48         # proj (48 EPI_RESID launches per forward: proj, fc2 alternate)
cap attention 'attention_kernel' 24
cap conv64 'conv64_kernel<\(int\)64,' 22          # a level-3 3x3 conv of the second forward
cap neckout 'conv64_kernel<\(int\)32,' 1
cap headout 'conv64_kernel<\(int\)16,' 3
cap convh128 'convh_kernel<\(int\)128,' 10
cap convh256 'convh_kernel<\(int\)256,' 10
cap convT 'umma_kernel<\(int\)256, \(int\)1, \(int\)4, \(bool\)0, \(int\)17>' 12
for f in gpurun_out/${TAG}_*.ncu-rep; do python tools/ncu_summary.py $f > ${f%.ncu-rep}.txt 2>/dev/null; done
ls gpurun_out/${TAG}_*.txt
