"""-m gpu: focal/shift recovery and post-processing kernels against the reference goldens and the oracle port."""
import os

import pytest
import torch

from moge_b200 import capi
from moge_b200.synthetic import synthetic_point_map
from oracle import moge_port
from gpu_util import rel_l2, stream

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _recover(points, mask, focal_in=None):
    B, H, W, _ = points.shape
    p = points.to(DEV).contiguous()
    m = mask.to(DEV).to(torch.uint8).contiguous() if mask is not None else None
    fin = focal_in.to(DEV).float().contiguous() if focal_in is not None else None
    f = torch.empty(B, device=DEV); s = torch.empty(B, device=DEV)
    capi.check(capi.lib().moge_recover_focal_shift(p.data_ptr(), None, capi.ptr(m), B, H, W, capi.ptr(fin), f.data_ptr(), s.data_ptr(), stream()))
    torch.cuda.synchronize()
    return f.cpu(), s.cpu()


def test_focal_shift_matches_reference_golden(golden_dir):
    cases = torch.load(os.path.join(golden_dir, "recover_focal_shift.pt"), weights_only=False)
    for i, c in enumerate(cases):
        pts, mask = synthetic_point_map(*c["args"][:7], seed=c["args"][7])
        f, s = _recover(pts, mask, c["focal_in"])
        # the kernel restates MINPACK lmdif (incl. its ftol=1e-3 early stop), so it lands on the reference's own iterate:
        # 1e-5 relative (SURVEY.md 8c (2) allowed 1e-4 for a merely converged solver)
        assert torch.allclose(f, c["focal"], rtol=1e-5, atol=1e-6), (i, f, c["focal"])
        assert torch.allclose(s, c["shift"], rtol=1e-5, atol=1e-6), (i, s, c["shift"])


def test_focal_shift_recovers_ground_truth():
    pts, mask = synthetic_point_map(3, 200, 300, 1.25, 0.4, 0.0, "all", seed=3)
    f, s = _recover(pts, mask)
    assert torch.allclose(f, torch.full((3,), 1.25), rtol=1e-3)         # LM stops at ftol=1e-3 like the reference
    assert torch.allclose(s, torch.full((3,), 0.4), rtol=3e-3, atol=1e-3)


@pytest.mark.parametrize("force_projection,apply_mask", [(True, True), (False, True), (True, False)])
def test_postprocess_matches_port(force_projection, apply_mask):
    B, H, W = 2, 90, 130
    pts, mask = synthetic_point_map(B, H, W, 0.9, 0.3, 0.01, "random", seed=9)
    g = torch.Generator().manual_seed(0)
    normal = torch.nn.functional.normalize(torch.randn(B, H, W, 3, generator=g), dim=-1)
    prob = torch.where(mask, 0.9, 0.1) + 0.05 * torch.rand(B, H, W, generator=g)
    scale = torch.tensor([1.7, 0.6])
    ref = moge_port.postprocess(pts, normal, prob, scale, W / H, force_projection=force_projection, apply_mask=apply_mask)
    focal, shift = moge_port.recover_focal_shift(pts, prob > 0.5)
    p = pts.to(DEV).contiguous(); n = normal.to(DEV).contiguous(); pr = prob.to(DEV).contiguous(); sc = scale.to(DEV)
    fo, sh = focal.to(DEV), shift.to(DEV)
    depth = torch.empty(B, H, W, device=DEV); nout = torch.empty_like(n)
    mout = torch.empty(B, H, W, dtype=torch.uint8, device=DEV); K = torch.empty(B, 3, 3, device=DEV)
    capi.check(capi.lib().moge_postprocess(p.data_ptr(), n.data_ptr(), pr.data_ptr(), sc.data_ptr(), fo.data_ptr(), sh.data_ptr(),
                                           B, H, W, int(force_projection), int(apply_mask), depth.data_ptr(), nout.data_ptr(),
                                           mout.data_ptr(), K.data_ptr(), stream()))
    torch.cuda.synchronize()
    m = ref["mask"]
    assert torch.equal(mout.cpu().bool(), m)
    assert rel_l2(K, ref["intrinsics"]) < 1e-6
    sel = m if apply_mask else torch.ones_like(m)
    assert rel_l2(p.cpu()[sel], ref["points"][sel]) < 1e-5
    assert rel_l2(depth.cpu()[sel], ref["depth"][sel]) < 1e-5
    assert rel_l2(nout.cpu(), ref["normal"]) < 1e-6
    if apply_mask:
        assert torch.isinf(p.cpu()[~m]).all() and torch.isinf(depth.cpu()[~m]).all()
