"""Wait-time attribution of conv64_kernel (needs a build with MG_EXTRA_FLAGS=-DMG_C64_DEBUG)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge_b200 import capi
L = capi.lib(); dev = "cuda"
B = int(os.environ.get("B", 32))
def counters(reset=1):
    out = (C.c_ulonglong * 8)()
    L.mg_debug_c64(out, reset)
    return list(out)
for (H, W, Cc, skip, both) in [(296, 296, 64, False, False), (296, 296, 64, True, True)]:
    Hp, Wp = H + 2, W + 2
    x = torch.randn(B, Hp, Wp, Cc, device=dev).half()
    w = (torch.randn(Cc, Cc, 3, 3, device=dev) / (9 * Cc) ** 0.5)
    b = torch.randn(Cc, device=dev)
    sk = torch.randn(B, Hp, Wp, Cc, device=dev).half() if skip else None
    o0 = torch.empty(B, Hp, Wp, Cc, device=dev, dtype=torch.float16)
    o1 = torch.empty(B, Hp, Wp, Cc, device=dev, dtype=torch.float16) if both else None
    st = capi.current_stream()
    f = lambda: capi.check(L.moge_op_conv(x.data_ptr(), w.data_ptr(), b.data_ptr(), capi.ptr(sk), capi.ptr(o0) if (both or not skip) else None, capi.ptr(o1) if both else (None if not skip else capi.ptr(o0)), B, H, W, Cc, Cc, 9, 0, capi.F16, st))
    f(); f(); counters(1)
    f()
    c = counters(1)
    n = 148.0
    tiles = B * ((H + 7) // 8) * ((W + 15) // 16) / n
    life = c[5] / n
    print(f"conv64 {H}x{W} skip={skip} both={both}: CTA lifetime {life:.0f} cyc, {tiles:.1f} tiles/CTA -> {life/tiles:.0f} cyc/tile")
    for name, v in zip(["producer waits free stage", "MMA waits data", "MMA waits free accumulator", "epilogue(w2) waits accumulator", "epilogue(w2) busy"], c[:5]):
        print(f"   {name:32s} {v/n:10.0f} cyc  = {v/n/life*100:5.1f} % of lifetime")
