import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moge.model.v2 import MoGeModel
from moge_b200 import capi
from moge_b200.configs import model_config
from moge_b200.synthetic import make_state_dict, synthetic_images
from oracle import moge_port
cfg = model_config("vitl", True); sd = make_state_dict(cfg, 5)
m = MoGeModel(**cfg); m.load_state_dict(sd); m = m.to("cuda").eval()
img = synthetic_images(1, 112, 140, 5).cuda()
out = m.forward(img, 120); torch.cuda.synchronize()
pts = out["points"].contiguous(); prob = out["mask"].contiguous()
f = torch.empty(1, device="cuda"); s = torch.empty(1, device="cuda")
capi.check(capi.lib().moge_recover_focal_shift(pts.data_ptr(), prob.data_ptr(), None, 1, 112, 140, None, f.data_ptr(), s.data_ptr(), capi.current_stream()))
torch.cuda.synchronize()
fp, sp = moge_port.recover_focal_shift(pts.cpu(), prob.cpu() > 0.5)
print("gpu focal/shift", f.item(), s.item(), " scipy", fp.item(), sp.item(), " rel diff", abs(f.item()-fp.item())/abs(fp.item()), abs(s.item()-sp.item())/abs(sp.item()))
z = pts[..., 2].cpu()
print("z min/median/max", z.min().item(), z.median().item(), z.max().item(), " n(z+s<0)", int((z + sp < 0).sum()), int((z + s.cpu() < 0).sum()))
import numpy as np, torch.nn.functional as F
uv = moge_port.view_plane_uv(140, 112, 140 / 112)
p_lr = F.interpolate(pts.cpu().permute(0, 3, 1, 2), (64, 64), mode="nearest").permute(0, 2, 3, 1).numpy()
uv_lr = F.interpolate(uv.permute(2, 0, 1)[None], (64, 64), mode="nearest")[0].permute(1, 2, 0).numpy()
m_lr = (F.interpolate((prob.cpu() > 0.5).float()[:, None], (64, 64), mode="nearest")[:, 0] > 0).numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/lm_case.npz", uv=uv_lr[m_lr[0]], p=p_lr[0][m_lr[0]])
