"""ncu targets for the decoder kernels through the C ABI: WHAT in {convT, conv64, conv64skip}. One warm call + one profiled."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge_b200 import capi
L = capi.lib(); dev = "cuda"; B = 32
what = os.environ.get("WHAT", "convT")
st = capi.current_stream()
def pad(B, H, W, C): return torch.randn(B, H + 2, W + 2, C, device=dev).half()
if what == "convT":
    H = W = 148; Cin, Cout = 128, 64
    x = pad(B, H, W, Cin); w = torch.randn(Cin, Cout, 2, 2, device=dev) / Cin ** 0.5; b = torch.randn(Cout, device=dev)
    o = torch.empty(B, 2 * H + 2, 2 * W + 2, Cout, device=dev, dtype=torch.float16)
    f = lambda: capi.check(L.moge_op_conv(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, o.data_ptr(), None, B, H, W, Cin, Cout, 1, 1, capi.F16, st))
else:
    H = W = 296; C = 64
    x = pad(B, H, W, C); w = torch.randn(C, C, 3, 3, device=dev) / (9 * C) ** 0.5; b = torch.randn(C, device=dev)
    sk = pad(B, H, W, C) if what == "conv64skip" else None
    o = torch.empty(B, H + 2, W + 2, C, device=dev, dtype=torch.float16)
    o2 = torch.empty(B, H + 2, W + 2, C, device=dev, dtype=torch.float16) if what == "conv64skip" else None
    f = lambda: capi.check(L.moge_op_conv(x.data_ptr(), w.data_ptr(), b.data_ptr(), capi.ptr(sk), o.data_ptr(), capi.ptr(o2), B, H, W, C, C, 9, 0, capi.F16, st))
f(); torch.cuda.synchronize(); f(); torch.cuda.synchronize()
print("done", what)
