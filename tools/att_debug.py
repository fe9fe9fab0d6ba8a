"""Wait-time attribution of attention_kernel (needs a build with MG_EXTRA_FLAGS=-DMG_ATT_DEBUG)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge_b200 import capi
L = capi.lib(); dev = "cuda"
B = int(os.environ.get("B", 32)); N = int(os.environ.get("N", 1370)); heads = 16; D = heads * 64
def counters(reset=1):
    out = (C.c_ulonglong * 8)()
    L.mg_debug_att(out, reset)
    return list(out)
qkv = torch.randn(B, N, 3 * D, device=dev).half()
out = torch.empty(B, N, D, device=dev, dtype=torch.float16)
st = capi.current_stream()
f = lambda: capi.check(L.moge_op_attention(qkv.data_ptr(), out.data_ptr(), B, N, D, heads, capi.F16, st))
f(); f(); counters(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); f(); e1.record(); torch.cuda.synchronize()
c = counters(1)
nq = (N + 255) // 256
items = B * heads * nq
nkv = (N + 127) // 128
ctas = min(items, torch.cuda.get_device_properties(0).multi_processor_count)
life = c[5] / ctas
tiles = items * nkv / ctas          # kv-tile iterations of one softmax group per CTA (group 0 takes part in every item)
print(f"attention B={B} N={N}: {e0.elapsed_time(e1):.3f} ms, {items} items on {ctas} persistent CTAs, lifetime {life:.0f} cyc/CTA, "
      f"{tiles:.0f} kv tiles per CTA -> {life / tiles:.0f} cyc per kv tile (pair of query tiles)")
for name, v in zip(["softmax(w4) waits S", "softmax(w4) waits PV(j-1)", "softmax(w4) kv loops total", "MMA waits P / S-free", "MMA waits Q/K/V",
                    "lifetime", "-", "softmax(w4) epilogues (wait last PV + O norm + store)"], c):
    print(f"   {name:56s} {v/ctas:12.0f} cyc  = {v/ctas/life*100:5.1f} % of lifetime   {v/ctas/tiles:8.0f} cyc / kv tile")
