"""Debug aid (GPU box): where does the end-to-end infer() deviation of a well-posed case come from -- the forward outputs or the
focal/shift solver?  python tools/wp_debug.py [seed H W tokens]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moge.model.v2 import MoGeModel
from moge_b200 import capi
from moge_b200.configs import model_config
from moge_b200.synthetic import make_state_dict, synthetic_images
from oracle import moge_port

seed, H, W, nt = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (13, 518, 1036, 700)))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = model_config("vitl", True)
sd = make_state_dict(cfg, seed, well_posed=True)
img = synthetic_images(1, H, W, seed)
m = MoGeModel(**cfg); m.load_state_dict(sd); m = m.to("cuda").eval()
fwd = m.forward(img.cuda(), nt)
torch.cuda.synchronize()
sdd = {k: v.cuda() for k, v in sd.items()}
ref = moge_port.forward(cfg, sdd, img.cuda(), nt)


def solve_engine(points, mask):
    f = torch.empty(1, device="cuda"); s = torch.empty(1, device="cuda")
    capi.check(capi.lib().moge_recover_focal_shift(points.contiguous().data_ptr(), mask.contiguous().data_ptr(), None, 1, H, W, None,
                                                   f.data_ptr(), s.data_ptr(), capi.current_stream()))
    torch.cuda.synchronize()
    return f.item(), s.item()


for name, src in (("engine forward", fwd), ("oracle fp32 forward", ref)):
    fe, se = solve_engine(src["points"].float(), src["mask"].float())
    fp, sp = moge_port.recover_focal_shift(src["points"].float().cpu(), src["mask"].float().cpu() > 0.5)
    print(f"{name:22s}: engine kernel (f, s) = ({fe:.6f}, {se:.6f})   SciPy (f, s) = ({fp.item():.6f}, {sp.item():.6f})")
d = (fwd["points"] - ref["points"]).cpu()
r = ref["points"].cpu()
print("forward points rel-L2", float(d.norm() / r.norm()), " z: mean signed rel err", float((d[..., 2] / r[..., 2]).mean()),
      " rms", float((d[..., 2] / r[..., 2]).pow(2).mean().sqrt()), " xy rms rel", float((d[..., :2].norm(dim=-1) / r[..., :2].norm(dim=-1).clamp_min(1e-6)).pow(2).mean().sqrt()))
