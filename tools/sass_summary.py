"""Instruction counts per kernel family of the shipped library (build container; needs cuobjdump, c++filt).
usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "moge_b200", "_lib", "libmoge_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2"]
agg, tot, nfun = collections.OrderedDict(), collections.Counter(), 0
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    base = re.sub(r"[<(].*", "", dem).replace("void ", "").replace("mg::", "")
    c = collections.Counter({k: len(re.findall(r"\b" + re.escape(k), part)) for k in KEYS})
    c["2CTA"] = len(re.findall(r"UTCHMMA\.2CTA", part))
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1; a[1].update(c); tot.update(c); nfun += 1
print("cuobjdump -sass moge_b200/_lib/libmoge_b200.so (final build): instruction counts per kernel family, template instantiations summed")
print("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st (tensor memory),")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier operations\n")
hdr = f"{'kernel family':20s} {'inst.':>5s} {'UTCHMMA':>8s} {'.2CTA':>6s} {'UTMALDG':>8s} {'UTMASTG':>8s} {'LDTM':>6s} {'STTM':>6s} {'UTCBAR':>7s} {'SYNCS':>6s} {'MUFU.EX2':>9s} {'FFMA2':>6s}"
print(hdr)
row = lambda k, n, c: f"{k:20s} {n:5d} {c['UTCHMMA']:8d} {c['2CTA']:6d} {c['UTMALDG']:8d} {c['UTMASTG']:8d} {c['LDTM']:6d} {c['STTM']:6d} {c['UTCBAR']:7d} {c['SYNCS']:6d} {c['MUFU.EX2']:9d} {c['FFMA2']:6d}"
for k, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["UTCHMMA"]):
    print(row(k, n, c))
print("\n" + row("total", nfun, tot))
print("\nNo UTMASTG: every global store is a 16-byte register store (full 128-byte rows after a swizzled shared-memory transpose);")
print("TMA is used for loads only (operands, halo boxes, weights, the staged fp32 residual of the proj GEMM).")
