"""CPU oracle for the MoGe-2 `forward()` / `infer()` hot path.  TEST INFRASTRUCTURE ONLY.

This is a functional fp32 restatement (plain torch ops on a state dict; no nn.Module, no CUDA
extension) of the reference algorithm.  It exists because `/root/reference` cannot travel to the GPU
box: it is pinned HERE against the unmodified reference (oracle/make_golden.py asserts agreement and
writes tests/golden/*.pt), and is then used there as the checker and as the timed CPU baseline.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Reference lines each function follows (paths relative to /root/reference):
  encode()            moge/model/modules.py:120-136, moge/model/dinov2/models/vision_transformer.py:187-243,283-333,
                      dinov2/layers/{patch_embed.py:68-81, block.py:88-113, attention.py:70-81, mlp.py:34-40, layer_scale.py:26-27}
  conv_stack()        moge/model/modules.py:47-68,155-165,242-254
  forward()           moge/model/v2.py:138-192, moge/utils/geometry_torch.py:40-52
  recover_focal_shift moge/utils/geometry_torch.py:115-170, moge/utils/geometry_numpy.py:79-112 (SciPy MINPACK 'lm')
  infer()             moge/model/v2.py:194-303 (+ utils3d@3fab839f intrinsics_from_focal_center / depth_map_to_point_map)

ATen operators (F.interpolate, F.layer_norm, F.conv2d, SDPA, GELU) and SciPy's least_squares are the
real third-party implementations, exactly the ones the reference dispatches to.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from moge_b200.configs import (backbone_dims, token_grid, default_num_tokens, PATCH, POS_GRID, INTERP_OFFSET,
                               LN_EPS)


# ----------------------------------------------------------------------------------------------- encoder
def pos_embed_for_grid(pos_embed: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """(1, 1+37*37, D) -> (1, 1+h*w, D).  vision_transformer.py:187-221 (non-onnx branch)."""
    n = pos_embed.shape[1] - 1
    if h * w == n and h == w:
        return pos_embed
    pe = pos_embed.float()
    M = int(math.sqrt(n))
    D = pe.shape[-1]
    grid = pe[:, 1:].reshape(1, M, M, D).permute(0, 3, 1, 2)
    sy, sx = float(h + INTERP_OFFSET) / M, float(w + INTERP_OFFSET) / M
    grid = F.interpolate(grid, mode="bicubic", antialias=False, scale_factor=(sy, sx))
    assert grid.shape[-2:] == (h, w)
    grid = grid.permute(0, 2, 3, 1).reshape(1, h * w, D)
    return torch.cat([pe[:, :1], grid], dim=1)


def vit_block(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, heads: int) -> torch.Tensor:
    B, N, D = x.shape
    y = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, heads, D // heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
    a = F.scaled_dot_product_attention(q, k, v)                       # scale = hd^-0.5
    a = a.transpose(1, 2).reshape(B, N, D)
    a = F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = x + sd[p + "ls1.gamma"] * a
    y = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    y = F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + sd[p + "ls2.gamma"] * y


def encode(cfg: Dict, sd: Dict[str, torch.Tensor], image: torch.Tensor, h: int, w: int):
    """image (B,3,H,W) in [0,1] -> (features (B,C0,h,w), cls (B,D)); modules.py:120-136."""
    D, depth, heads = backbone_dims(cfg["encoder"]["backbone"])
    taps = list(cfg["encoder"]["intermediate_layers"])
    bb = "encoder.backbone."
    x = F.interpolate(image, (h * PATCH, w * PATCH), mode="bilinear", align_corners=False, antialias=True)
    x = (x - sd["encoder.image_mean"]) / sd["encoder.image_std"]
    x = F.conv2d(x, sd[bb + "patch_embed.proj.weight"], sd[bb + "patch_embed.proj.bias"], stride=PATCH)
    B = x.shape[0]
    x = x.flatten(2).transpose(1, 2)                                   # (B, T, D)
    x = torch.cat([sd[bb + "cls_token"].expand(B, -1, -1), x], dim=1)
    x = x + pos_embed_for_grid(sd[bb + "pos_embed"], h, w).to(x.dtype)
    feats = None
    cls = None
    j = 0
    for i in range(depth):
        x = vit_block(x, sd, f"{bb}blocks.{i}.", heads)
        if i in taps:
            y = F.layer_norm(x, (D,), sd[bb + "norm.weight"], sd[bb + "norm.bias"], LN_EPS)
            cls = y[:, 0]
            f = y[:, 1:].permute(0, 2, 1).reshape(B, D, h, w)
            f = F.conv2d(f, sd[f"encoder.output_projections.{j}.weight"], sd[f"encoder.output_projections.{j}.bias"])
            feats = f if feats is None else feats + f
            j += 1
    return feats, cls


# ----------------------------------------------------------------------------------------------- decoder
def view_plane_uv(width: int, height: int, aspect: float, dtype=torch.float32, device=None) -> torch.Tensor:
    """(H, W, 2) pixel-centre UV scaled to the unit-diagonal view plane; geometry_torch.py:40-52."""
    sx = aspect / (1 + aspect ** 2) ** 0.5
    sy = 1 / (1 + aspect ** 2) ** 0.5
    u = torch.linspace(-sx * (width - 1) / width, sx * (width - 1) / width, width, dtype=dtype, device=device)
    v = torch.linspace(-sy * (height - 1) / height, sy * (height - 1) / height, height, dtype=dtype, device=device)
    uu, vv = torch.meshgrid(u, v, indexing="xy")
    return torch.stack([uu, vv], dim=-1)


def _conv3(x, w, b):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w, b)


def conv_stack(ccfg: Dict, sd: Dict[str, torch.Tensor], name: str, inputs):
    """modules.py:242-254 with Resampler :155-165 and ResidualConvBlock :47-68 (norms = Identity)."""
    widths = ccfg["dim_res_blocks"]
    nlev = len(widths)
    nres = ccfg["num_res_blocks"] if isinstance(ccfg["num_res_blocks"], (list, tuple)) else [ccfg["num_res_blocks"]] * nlev
    res = ccfg["resamplers"] if isinstance(ccfg["resamplers"], (list, tuple)) else [ccfg["resamplers"]] * (nlev - 1)
    if ccfg.get("res_block_in_norm", "layer_norm") != "none" or ccfg.get("res_block_hidden_norm", "group_norm") != "none":
        raise NotImplementedError("MoGe-2 decoders use norm-free residual blocks")
    outs = []
    x = None
    for l in range(nlev):
        key = f"{name}.input_blocks.{l}."
        f = F.conv2d(inputs[l], sd[key + "weight"], sd[key + "bias"]) if key + "weight" in sd else inputs[l]
        x = f if l == 0 else x + f
        for r in range(nres[l]):
            p = f"{name}.res_blocks.{l}.{r}.layers."
            y = _conv3(F.relu(x), sd[p + "2.weight"], sd[p + "2.bias"])
            y = _conv3(F.relu(y), sd[p + "5.weight"], sd[p + "5.bias"])
            x = x + y
        key = f"{name}.output_blocks.{l}."
        outs.append(F.conv2d(x, sd[key + "weight"], sd[key + "bias"]) if key + "weight" in sd else x)
        if l < nlev - 1:
            p = f"{name}.resamplers.{l}."
            if res[l] == "conv_transpose":
                x = F.conv_transpose2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=2)
            elif res[l] == "bilinear":
                x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
            else:
                raise NotImplementedError(res[l])
            x = _conv3(x, sd[p + "1.weight"], sd[p + "1.bias"])
    return outs


def remap_points(points: torch.Tensor, mode: str) -> torch.Tensor:
    """v2.py:122-136"""
    if mode == "linear":
        return points
    if mode == "sinh":
        return torch.sinh(points)
    xy, z = points[..., :2], points[..., 2:]
    if mode == "exp":
        z = torch.exp(z)
        return torch.cat([xy * z, z], dim=-1)
    if mode == "sinh_exp":
        return torch.cat([torch.sinh(xy), torch.exp(z)], dim=-1)
    raise ValueError(f"Invalid remap output type: {mode}")


@torch.no_grad()
def forward(cfg: Dict, sd: Dict[str, torch.Tensor], image: torch.Tensor, num_tokens: int) -> Dict[str, torch.Tensor]:
    """v2.py:138-192.  image (B,3,H,W); runs in the dtype and on the device of `image` / `sd` (fp32 on CPU = the oracle;
    on 'cuda' it is the same-box PyTorch comparator and the large-shape checker of the -m gpu tests)."""
    B, _, H, W = image.shape
    aspect = W / H
    h, w = token_grid(H, W, int(num_tokens))
    feat, cls = encode(cfg, sd, image, h, w)
    inputs = []
    for l in range(5):
        uv = view_plane_uv(w * 2 ** l, h * 2 ** l, aspect, image.dtype, image.device).permute(2, 0, 1)[None].expand(B, -1, -1, -1)
        inputs.append(torch.cat([feat, uv], dim=1) if l == 0 else uv)
    neck = conv_stack(cfg["neck"], sd, "neck", inputs)
    out = {}
    for head in ("points_head", "normal_head", "mask_head"):
        if cfg.get(head) is not None:
            y = conv_stack(cfg[head], sd, head, neck)[-1]
            out[head] = F.interpolate(y, (H, W), mode="bilinear", align_corners=False, antialias=False)
    ret = {}
    if "points_head" in out:
        ret["points"] = remap_points(out["points_head"].permute(0, 2, 3, 1), cfg.get("remap_output", "linear"))
    if "normal_head" in out:
        ret["normal"] = F.normalize(out["normal_head"].permute(0, 2, 3, 1), dim=-1)
    if "mask_head" in out:
        ret["mask"] = out["mask_head"].squeeze(1).sigmoid()
    if cfg.get("scale_head") is not None:
        y = cls
        nlin = len(cfg["scale_head"]["dims"]) - 1
        for i in range(nlin):
            y = F.linear(y, sd[f"scale_head.{2 * i}.weight"], sd[f"scale_head.{2 * i}.bias"])
            if i < nlin - 1:
                y = F.relu(y)
        ret["metric_scale"] = y.squeeze(1).exp()
    return ret


# --------------------------------------------------------------------------------- focal / shift recovery
def _lm_shift(uv: np.ndarray, xyz: np.ndarray, focal: Optional[float]):
    """geometry_numpy.py:79-112: 1-D MINPACK LM on the z-shift, focal in closed form (or given)."""
    from scipy.optimize import least_squares
    uv = uv.reshape(-1, 2)
    xy, z = xyz[..., :2].reshape(-1, 2), xyz[..., 2].reshape(-1)

    def residual(shift):
        proj = xy / (z + shift)[:, None]
        f = (proj * uv).sum() / np.square(proj).sum() if focal is None else focal
        return (f * proj - uv).ravel()

    sol = least_squares(residual, x0=0, ftol=1e-3, method="lm")
    shift = sol["x"].squeeze().astype(np.float32)
    if focal is None:
        proj = xy / (z + shift)[:, None]
        return shift, (proj * uv).sum() / np.square(proj).sum()
    return shift, focal


def recover_focal_shift(points: torch.Tensor, mask: Optional[torch.Tensor] = None,
                        focal: Optional[torch.Tensor] = None, size=(64, 64)):
    """geometry_torch.py:115-170.  points (B,H,W,3), mask (B,H,W) bool -> focal (B,), shift (B,)."""
    dev_in = points.device
    points = points.detach().cpu().float()
    mask = None if mask is None else mask.detach().cpu()
    focal = None if focal is None else focal.detach().cpu()
    B, H, W, _ = points.shape
    uv = view_plane_uv(W, H, W / H, points.dtype)
    p_lr = F.interpolate(points.permute(0, 3, 1, 2), size, mode="nearest").permute(0, 2, 3, 1).numpy()
    uv_lr = F.interpolate(uv.permute(2, 0, 1)[None], size, mode="nearest")[0].permute(1, 2, 0).numpy()
    m_lr = None if mask is None else (F.interpolate(mask.float()[:, None], size, mode="nearest")[:, 0] > 0).numpy()
    fs, ss = [], []
    for i in range(B):
        pi = p_lr[i] if m_lr is None else p_lr[i][m_lr[i]]
        ui = uv_lr if m_lr is None else uv_lr[m_lr[i]]
        if ui.reshape(-1, 2).shape[0] < 2:
            fs.append(1.0 if focal is None else float(focal[i])); ss.append(0.0)
            continue
        s, f = _lm_shift(ui, pi, None if focal is None else float(focal[i]))
        fs.append(float(f)); ss.append(float(s))
    return torch.tensor(fs, dtype=points.dtype, device=dev_in), torch.tensor(ss, dtype=points.dtype, device=dev_in)


def postprocess(points: torch.Tensor, normal, mask, metric_scale, aspect: float, fov_x=None,
                force_projection: bool = True, apply_mask: bool = True, focal_shift=None) -> Dict[str, torch.Tensor]:
    """v2.py:246-298 on raw forward() outputs (all fp32).  Batched; no squeeze.
    `focal_shift=(focal, shift)` injects a given solution of the focal/shift solve (test cut point for the
    post-processing chain alone); by default it is solved here with SciPy like the reference."""
    points = points.clone().float()
    mask_b = None if mask is None else mask.float() > 0.5
    if focal_shift is not None:
        focal, shift = focal_shift
    elif fov_x is None:
        focal, shift = recover_focal_shift(points, mask_b)
    else:
        focal = aspect / (1 + aspect ** 2) ** 0.5 / torch.tan(torch.deg2rad(torch.as_tensor(fov_x, dtype=points.dtype, device=points.device) / 2))
        if focal.ndim == 0:
            focal = focal[None].expand(points.shape[0])
        _, shift = recover_focal_shift(points, mask_b, focal=focal)
    fx = focal / 2 * (1 + aspect ** 2) ** 0.5 / aspect
    fy = focal / 2 * (1 + aspect ** 2) ** 0.5
    K = torch.zeros(points.shape[0], 3, 3, dtype=points.dtype, device=points.device)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = fx, fy, 0.5, 0.5, 1.0
    points[..., 2] += shift[:, None, None]
    if mask_b is not None:
        mask_b = mask_b & (points[..., 2] > 0)
    depth = points[..., 2].clone()
    if force_projection:
        Hh, Ww = depth.shape[-2:]
        u = (torch.arange(Ww, dtype=depth.dtype, device=depth.device) + 0.5) / Ww
        v = (torch.arange(Hh, dtype=depth.dtype, device=depth.device) + 0.5) / Hh
        x = (u[None, None, :] - 0.5) / fx[:, None, None] * depth
        y = (v[None, :, None] - 0.5) / fy[:, None, None] * depth
        points = torch.stack([x, y, depth], dim=-1)
    if metric_scale is not None:
        points = points * metric_scale[:, None, None, None]
        depth = depth * metric_scale[:, None, None]
    if apply_mask and mask_b is not None:
        points = torch.where(mask_b[..., None], points, torch.inf)
        depth = torch.where(mask_b, depth, torch.inf)
        if normal is not None:
            normal = torch.where(mask_b[..., None], normal, torch.zeros_like(normal))
    ret = {"points": points, "intrinsics": K, "depth": depth}
    if mask_b is not None:
        ret["mask"] = mask_b
    if normal is not None:
        ret["normal"] = normal
    return ret


@torch.no_grad()
def infer(cfg: Dict, sd: Dict[str, torch.Tensor], image: torch.Tensor, num_tokens: Optional[int] = None,
          resolution_level: int = 9, force_projection: bool = True, apply_mask: bool = True, fov_x=None):
    """v2.py:194-303 with use_fp16=False.  image (B,3,H,W) or (3,H,W) fp32 in [0,1]."""
    squeeze = image.dim() == 3
    if squeeze:
        image = image[None]
    H, W = image.shape[-2:]
    if num_tokens is None:
        num_tokens = default_num_tokens(cfg.get("num_tokens_range", [1200, 3600]), resolution_level)
    raw = forward(cfg, sd, image.float(), num_tokens)
    ret = postprocess(raw.get("points"), raw.get("normal"), raw.get("mask"), raw.get("metric_scale"), W / H,
                      fov_x=fov_x, force_projection=force_projection, apply_mask=apply_mask)
    if squeeze:
        ret = {k: v[0] for k, v in ret.items()}
    return ret
