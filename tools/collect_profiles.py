"""Copy the evidence of a tools/final_validate.sh run from gpurun_out/ (scratch) into profiles/ (tracked) and derive
profiles/<TAG>_ncu_traffic.json (DRAM bytes per launch of the dominant kernel of each class, next to its algorithmic bytes).
usage (build container, after the GPU call): python tools/collect_profiles.py r2"""
import json, os, re, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
names = ["bench_n1.json", "bench_reference_arm.json", "bench_config3.json", "bench_config5.json", "bench_n1_tokens3600_batch8.json",
         "ops_profile.json", "gpu_baseline_kernels.json", "error_attribution.json", "launch_list_ncu_summary.txt",
         "gemm_qkv.txt", "gemm_proj.txt", "attention.txt", "conv64.txt", "neckout.txt", "headout.txt", "convh128.txt", "convh256.txt", "convT.txt"]
for n in names:
    a = os.path.join(src, f"{tag}_{n}")
    if os.path.exists(a) and os.path.getsize(a) > 0:
        b = os.path.join(dst, f"{tag}_{'ncu_full_' if n.endswith('.txt') and 'launch_list' not in n else ''}{n}")
        shutil.copyfile(a, b)
        print("copied", os.path.basename(b))
    else:
        print("MISSING", a)


def dram_bytes(txt):
    m = re.search(r"traffic \(dram read\+write\)\s+([0-9.]+)\s+(\w+)", open(txt).read())
    if not m:
        return None
    v, u = float(m.group(1)), m.group(2).lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


ops = os.path.join(dst, f"{tag}_ops_profile.json")
traffic = {}
if os.path.exists(ops):
    prof = json.load(open(ops))

    def alg(prefix):
        for o in prof:
            if o["name"].startswith(prefix):
                return o["bytes"]
        return None
    for key, fname, opname, kernel in (("gemm", "gemm_qkv", "gemm.qkv", "umma2_kernel<EPI_STORE16> qkv M=43840 N=3072 K=1024"),
                                       ("attention", "attention", "attention", "attention_kernel B=32 N=1370 16 heads"),
                                       ("decoder", "conv64", "conv3x3.res_b.neck.l3", "conv64_kernel<64,EPI_DEC> 3x3 C=64 296x296")):
        t = os.path.join(dst, f"{tag}_ncu_full_{fname}.txt")
        if os.path.exists(t):
            d = dram_bytes(t)
            if d:
                traffic[key] = {"kernel": kernel, "dram_bytes": d, "algorithmic_bytes": alg(opname), "file": f"profiles/{tag}_ncu_full_{fname}.txt"}
    json.dump(traffic, open(os.path.join(dst, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    print("traffic", traffic)
