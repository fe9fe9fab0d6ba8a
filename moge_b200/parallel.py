"""Data-parallel plumbing for the MoGe-2 hot path: images are independent units, so ranks shard
the batch with NO collective on the data path.  NCCL (torch.distributed) is used for exactly two things:
one broadcast of the checkpoint tensors from rank 0 at load time, and the gather of the output maps to rank 0.
One process per GPU (torchrun); works with the gloo backend on CPU tensors for the host-logic tests."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `total` images for `rank` (first `total % world` ranks get one extra)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_state_dict(state: Optional[Dict[str, torch.Tensor]], device: torch.device, src: int = 0,
                         group=None) -> Dict[str, torch.Tensor]:
    """Rank `src` holds `state` (CPU or device tensors); every rank returns the same dict on `device`.
    Tensors travel as ONE flat fp32 buffer (one NCCL broadcast), after a small object broadcast of the key/shape list."""
    rank = dist.get_rank(group)
    meta = None
    if rank == src:
        assert state is not None
        keys = [k for k, v in state.items() if v.is_floating_point()]
        meta = [(k, tuple(state[k].shape)) for k in keys]
    box = [meta]
    dist.broadcast_object_list(box, src=src, group=group)
    meta = box[0]
    total = sum(int(torch.Size(s).numel()) for _, s in meta)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if rank == src:
        off = 0
        for k, s in meta:
            n = int(torch.Size(s).numel())
            flat[off:off + n].copy_(state[k].reshape(-1).to(torch.float32))
            off += n
    dist.broadcast(flat, src=src, group=group)
    out, off = {}, 0
    for k, s in meta:
        n = int(torch.Size(s).numel())
        out[k] = flat[off:off + n].view(s)
        off += n
    return out


def gather_outputs(local: Dict[str, torch.Tensor], counts: List[int], dst: int = 0, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Gather per-rank output dicts (leading dim = local batch) onto rank `dst`, concatenated in rank order.
    `counts[r]` = images held by rank r.  Uses point-to-point send/recv so only `dst` holds the full batch."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    keys = sorted(local.keys())
    if rank == dst:
        full = {}
        for k in keys:
            t = local[k]
            as_u8 = t.dtype == torch.bool
            parts = []
            for r in range(world):
                if r == dst:
                    parts.append(t.view(torch.uint8) if as_u8 else t)
                    continue
                buf = torch.empty((counts[r],) + tuple(t.shape[1:]), dtype=torch.uint8 if as_u8 else t.dtype, device=t.device)
                if counts[r]:
                    dist.recv(buf, src=r, group=group)
                parts.append(buf)
            cat = torch.cat(parts, dim=0)
            full[k] = cat.view(torch.bool) if as_u8 else cat
        return full
    for k in keys:
        t = local[k]
        if t.shape[0]:
            dist.send((t.view(torch.uint8) if t.dtype == torch.bool else t).contiguous(), dst=dst, group=group)
    return None
