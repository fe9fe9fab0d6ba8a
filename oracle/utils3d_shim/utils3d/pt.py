"""utils3d.pt subset: intrinsics_from_focal_center, depth_map_to_point_map (see package docstring)."""
import torch


def intrinsics_from_focal_center(fx, fy, cx, cy):
    fx, fy, cx, cy = torch.broadcast_tensors(*[torch.as_tensor(v) for v in (fx, fy, cx, cy)])
    fx = fx.to(torch.promote_types(fx.dtype, torch.float32)) if not fx.is_floating_point() else fx
    K = torch.zeros(*fx.shape, 3, 3, dtype=fx.dtype, device=fx.device)
    K[..., 0, 0] = fx
    K[..., 1, 1] = fy
    K[..., 0, 2] = cx.to(fx)
    K[..., 1, 2] = cy.to(fx)
    K[..., 2, 2] = 1
    return K


def depth_map_to_point_map(depth, intrinsics):
    """depth (..., H, W), normalized intrinsics (..., 3, 3) -> camera-space points (..., H, W, 3).
    Pixel-centre UV: u = (j + 0.5) / W, v = (i + 0.5) / H."""
    H, W = depth.shape[-2:]
    u = (torch.arange(W, dtype=depth.dtype, device=depth.device) + 0.5) / W
    v = (torch.arange(H, dtype=depth.dtype, device=depth.device) + 0.5) / H
    fx, fy = intrinsics[..., 0, 0], intrinsics[..., 1, 1]
    cx, cy = intrinsics[..., 0, 2], intrinsics[..., 1, 2]
    x = (u[None, :] - cx[..., None, None]) / fx[..., None, None] * depth
    y = (v[:, None] - cy[..., None, None]) / fy[..., None, None] * depth
    return torch.stack([x, y, depth], dim=-1)
