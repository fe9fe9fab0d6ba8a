"""Run the engine forward several times on the same input and report run-to-run differences (should be bit-exact)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge.model.v2 import MoGeModel
from moge_b200.configs import model_config
from moge_b200.synthetic import make_state_dict, synthetic_images
for size, shape, nt in [("vits", (2, 140, 98), 117), ("vitl", (1, 112, 140), 120), ("vits", (1, 224, 224), 3600)]:
    cfg = model_config(size, True); sd = make_state_dict(cfg, 1)
    m = MoGeModel(**cfg); m.load_state_dict(sd); m = m.to("cuda").eval()
    img = synthetic_images(*shape, 5).cuda()
    ref = None
    for it in range(6):
        out = m.forward(img, nt); torch.cuda.synchronize()
        out = {k: v.clone() for k, v in out.items()}
        if ref is None: ref = out; continue
        d = {k: float((out[k] - ref[k]).abs().max()) for k in out}
        nz = {k: int((out[k] != ref[k]).sum()) for k in out}
        print(size, shape, "run", it, "max|diff|", d, "n_diff", nz, flush=True)
    del m
