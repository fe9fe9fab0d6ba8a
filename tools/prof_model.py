"""ncu target: the real launch list of the benchmarked configuration (ViT-L fp16, B x 518x518, T=1369), two forward() calls.
Pick one launch of a kernel with   ncu --kernel-name-base demangled -k regex:<pattern> -s <skip> -c 1 ... python tools/prof_model.py
(first forward = launches 0..N-1 of each kernel; the second forward is the warm one: skip past the first)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge.model.v2 import MoGeModel
from moge_b200.configs import model_config
from moge_b200.synthetic import make_state_dict
B = int(os.environ.get("B", 32))
cfg = model_config(os.environ.get("SIZE", "vitl"), True)
m = MoGeModel(**cfg); m.load_state_dict(make_state_dict(cfg, 0)); m = m.to("cuda").eval()
x = torch.rand(B, 3, 518, 518, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    m.forward(x, 1369)
torch.cuda.synchronize()
print("done", [n for n, _, _ in m.engine_ops()][:3])
