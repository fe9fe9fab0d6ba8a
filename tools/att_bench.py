"""Stand-alone launches of attention_kernel on the benchmark shape (for ncu captures and quick timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge_b200 import capi
L = capi.lib()
B = int(os.environ.get("B", 32)); N = int(os.environ.get("N", 1370)); heads = 16; D = heads * 64
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 2).half()
out = torch.empty(B, N, D, device="cuda", dtype=torch.float16)
for _ in range(int(os.environ.get("REPS", 3))):
    capi.check(L.moge_op_attention(qkv.data_ptr(), out.data_ptr(), B, N, D, heads, capi.F16, capi.current_stream()))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
