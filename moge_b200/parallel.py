"""Data-parallel plumbing for the MoGe-2 hot path: images are independent units, so ranks shard
the batch with NO collective on the data path.  One process per GPU (torchrun).
  * `broadcast_state_dict`: one NCCL broadcast of the checkpoint tensors from rank 0 at load time;
  * `PeerGatherer`: the per-step gather of the output maps to rank 0 over NVLink peer memory (CUDA-IPC staging slots, copy-engine
    pulls, device-side flags; libmoge_b200's moge_peer_* entry points) -- the product path, no SM time;
  * `OutputGatherer` / `gather_outputs`: the same gather as NCCL (or gloo) send/recv -- kept for comparison in bench.py and for
    the world-size-2 gloo tests of the host logic on CPU tensors."""
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
    Tensors travel as ONE flat fp32 buffer (one NCCL broadcast), after a small object broadcast of the key/shape list.
    fp32 on purpose: the engine's load-time folds (LayerNorm into qkv / fc1, tap projections into the neck, the head tails) are
    computed from the fp32 originals before the single rounding to 16 bit, so every rank ends up with bit-identical packed weights
    and outputs; shipping 16-bit-rounded originals would round twice on the receiving ranks."""
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


class OutputGatherer:
    """Pipelined gather of per-rank `infer()` outputs onto rank `dst` (BASELINE.json configs[3]: "inputs resident on each GPU ->
    all outputs resident on rank 0").  `submit(local)` enqueues ONE grouped NCCL send/recv batch (every key, every peer) on a side
    stream, straight into preallocated per-key buffers of the full batch on `dst` -- no per-key serial send/recv, no torch.cat --
    so the transfer of step i runs under the compute of step i+1.  `fence()` makes the current stream wait for everything
    submitted.  With the gloo backend (CPU tensors, host-logic tests) the same ops run synchronously."""

    def __init__(self, counts: List[int], dst: int = 0, group=None, device: Optional[torch.device] = None, slots: int = 2):
        self.counts, self.dst, self.group = list(counts), dst, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.offsets = [sum(self.counts[:r]) for r in range(self.world)]
        self.device = device
        self.cuda = device is not None and device.type == "cuda"
        self.slots = slots
        self.full: List[Optional[Dict[str, torch.Tensor]]] = [None] * slots
        self.n = 0
        if self.cuda:
            with torch.cuda.device(device):
                self.side = torch.cuda.Stream()
                self.done = [torch.cuda.Event() for _ in range(slots)]

    def _buffers(self, slot: int, local: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self.full[slot] is None:
            total = sum(self.counts)
            self.full[slot] = {k: torch.empty((total,) + tuple(t.shape[1:]), dtype=torch.uint8 if t.dtype == torch.bool else t.dtype,
                                              device=t.device) for k, t in local.items()}
        return self.full[slot]

    def _post(self, local: Dict[str, torch.Tensor], full: Optional[Dict[str, torch.Tensor]]):
        keys = sorted(local.keys())
        ops = []
        if self.rank == self.dst:
            o, n = self.offsets[self.rank], self.counts[self.rank]
            for k in keys:
                t = local[k]
                full[k][o:o + n].copy_(t.view(torch.uint8) if t.dtype == torch.bool else t)
                for r in range(self.world):
                    if r != self.dst and self.counts[r]:
                        ops.append(dist.P2POp(dist.irecv, full[k][self.offsets[r]:self.offsets[r] + self.counts[r]], r, self.group))
        elif self.counts[self.rank]:
            for k in keys:
                t = local[k]
                ops.append(dist.P2POp(dist.isend, (t.view(torch.uint8) if t.dtype == torch.bool else t).contiguous(), self.dst, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def submit(self, local: Dict[str, torch.Tensor]) -> Optional[Dict[str, torch.Tensor]]:
        """Queue the gather of one step's outputs.  Returns (on `dst`) the dict of full-batch buffers this step lands in -- valid
        after `fence()` / `wait()`; bool outputs come back as bool views."""
        slot = self.n % self.slots
        self.n += 1
        full = self._buffers(slot, local) if self.rank == self.dst else None
        if not self.cuda:
            for w in self._post(local, full):
                w.wait()
        else:
            with torch.cuda.device(self.device):
                ready = torch.cuda.Event()
                ready.record()                                  # the outputs are complete on the producing stream
                self.side.wait_event(ready)
                self.side.wait_event(self.done[slot])           # (no-op the first time) previous user of this slot's buffers
                with torch.cuda.stream(self.side):
                    for w in self._post(local, full):
                        w.wait()                                # stream-level wait: orders the side stream after the transfers
                    self.done[slot].record(self.side)
                for t in local.values():
                    t.record_stream(self.side)
        if full is None:
            return None
        return {k: (v.view(torch.bool) if local[k].dtype == torch.bool else v) for k, v in full.items()}

    def fence(self, stream=None) -> None:
        if not self.cuda:
            return
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream()
            for ev in self.done:
                st.wait_event(ev)

    def wait(self) -> None:
        if self.cuda:
            self.fence()
            torch.cuda.current_stream(self.device).synchronize()


class PeerSetupError(RuntimeError):
    """PeerGatherer could not allocate / map its CUDA-IPC staging memory on some rank; raised on EVERY rank of the group."""


class PeerGatherer:
    """The output gather of BASELINE.json configs[3] over NVLink peer memory, with ZERO SM time (libmoge_b200's moge_peer_* entry
    points): every rank copies its step outputs into an IPC-exported staging slot and raises a flag; rank `dst` waits for the flag
    (a one-thread poll kernel on a side stream), PULLS the slot with copy-engine DMAs straight into its preallocated full-batch
    buffers and raises a "consumed" flag the producers check before they reuse the slot.  Nothing blocks the host; the transfer of
    step i runs under the compute of step i+1 and, unlike a NCCL send/recv pair, without copy kernels competing for SMs with the
    engine's persistent one-CTA-per-SM kernels.  `torch.distributed` (any backend) is used once, to exchange the IPC handles.

        gat = PeerGatherer([B] * world, device)        # collective: every rank constructs it
        res = gat.submit(model.infer(x))               # every step; on `dst`: dict of full-batch buffers (valid after fence/wait)
        gat.fence()                                    # current stream waits for everything submitted so far
    """

    SLOTS = 2

    def __init__(self, counts: List[int], device: torch.device, dst: int = 0, group=None):
        import ctypes as C
        from . import capi
        self.C, self.capi, self.L = C, capi, capi.lib()
        self.counts, self.dst, self.group, self.device = list(counts), dst, group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.offsets = [sum(self.counts[:r]) for r in range(self.world)]
        self.n = 0
        self.layout = None               # [(key, dtype, per-image shape, byte offset in a slot)] fixed by the first submit
        self.slot_bytes = 0
        self.stage = None                # this rank's staging buffer (device pointer), SLOTS slots + flags at the end
        self.peers = {}                  # dst only: rank -> mapped staging pointer
        self.flags_dst = None            # producers: mapped pointer to dst's "consumed" flags
        self.full: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.SLOTS
        import os
        self.debug = os.environ.get("MOGE_B200_GATHER_DEBUG") == "1"       # CUDA-event timing of the staging copy / the pulls
        self.dbg_events = []
        with torch.cuda.device(device):
            self.side = [torch.cuda.Stream() for _ in range(max(1, self.world))]
            self.done = [torch.cuda.Event() for _ in range(self.SLOTS)]

    def debug_report(self):
        """(debug mode) mean milliseconds of the timed sections, by label"""
        torch.cuda.synchronize(self.device)
        acc = {}
        for label, e0, e1 in self.dbg_events:
            acc.setdefault(label, []).append(e0.elapsed_time(e1))
        return {k: sum(v) / len(v) for k, v in acc.items()}

    def _mark(self, label, stream):
        if not self.debug:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        return (label, e0)

    def _end(self, tok, stream):
        if tok is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(stream)
            self.dbg_events.append((tok[0], tok[1], e1))

    # -- one-time setup, at the first submit (the output shapes are known then)
    def _setup(self, local: Dict[str, torch.Tensor]):
        C, L, capi = self.C, self.L, self.capi
        keys = sorted(local.keys())
        off, layout = 0, []
        per_rank_max = max(self.counts)
        for k in keys:
            t = local[k]
            per_img = t.numel() // max(t.shape[0], 1) * t.element_size()
            layout.append((k, t.dtype, tuple(t.shape[1:]), off, per_img))
            off += (per_img * per_rank_max + 255) // 256 * 256
        self.layout, self.slot_bytes = layout, off
        flag_bytes = 256
        total = self.SLOTS * self.slot_bytes + flag_bytes
        # Every rank reaches the same verdict: a rank whose allocation / mapping failed reports it through the (host-side) object
        # exchange and ALL ranks raise PeerSetupError together, before any device-side flag is waited on -- the caller can then
        # fall back to OutputGatherer collectively instead of leaving the healthy ranks polling a flag nobody will raise.
        with torch.cuda.device(self.device):
            ptr = C.c_void_p()
            handle = C.create_string_buffer(64)
            err = None
            try:
                capi.check(L.moge_peer_alloc(total, C.byref(ptr), handle))
                self.stage = ptr.value
                self.flag_base = self.stage + self.SLOTS * self.slot_bytes    # int32 flags: [ready[s] for s] then [consumed[s] for s]
            except Exception as ex:                                            # noqa: BLE001 -- reported to every rank below
                err = f"rank {self.rank}: {ex}"
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (self.rank, bytes(handle.raw), total, err), group=self.group)
            errs = [g[3] for g in gathered if g[3]]
            if not errs:
                try:
                    if self.rank == self.dst:
                        for r, h, tot, _ in gathered:
                            if r == self.dst or not self.counts[r]:
                                continue
                            p = C.c_void_p()
                            capi.check(L.moge_peer_open(h, C.byref(p)))
                            self.peers[r] = p.value
                    else:
                        h = [g for g in gathered if g[0] == self.dst][0][1]
                        p = C.c_void_p()
                        capi.check(L.moge_peer_open(h, C.byref(p)))
                        self.flags_dst = p.value + self.SLOTS * self.slot_bytes
                except Exception as ex:                                        # noqa: BLE001
                    err = f"rank {self.rank}: {ex}"
                opened = [None] * self.world
                dist.all_gather_object(opened, err, group=self.group)
                errs = [e for e in opened if e]
            if errs:
                self.close()
                self.layout = None
                raise PeerSetupError("peer-memory gather unavailable: " + "; ".join(errs))
        dist.barrier(group=self.group)

    def _ready_flag(self, base, slot):
        return base + self.SLOTS * self.slot_bytes + 4 * slot

    def submit(self, local: Dict[str, torch.Tensor]) -> Optional[Dict[str, torch.Tensor]]:
        C, L, capi = self.C, self.L, self.capi
        if self.layout is None:
            self._setup(local)
        slot = self.n % self.SLOTS
        seq = self.n + 1                                   # flag value of this step
        self.n += 1
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            if self.rank != self.dst:
                if not self.counts[self.rank]:
                    return None
                st = cur.cuda_stream
                if seq > self.SLOTS:                       # the slot's previous contents must have been pulled by dst
                    capi.check(L.moge_peer_flag_wait(self.flags_dst + 4 * (self.SLOTS + slot), seq - self.SLOTS, st))
                base = self.stage + slot * self.slot_bytes
                tok = self._mark("producer: staging copy", cur)
                for k, dt, shp, off, per_img in self.layout:
                    t = local[k].contiguous()
                    capi.check(L.moge_peer_copy(base + off, t.data_ptr(), per_img * t.shape[0], st))
                self._end(tok, cur)
                capi.check(L.moge_peer_flag_set(self._ready_flag(self.stage, slot), seq, st))
                return None
            # ---- dst: pull every peer's slot with DMAs on side streams, own outputs by a local copy
            if self.full[slot] is None:
                total = sum(self.counts)
                self.full[slot] = {k: torch.empty((total,) + shp, dtype=torch.uint8 if dt == torch.bool else dt, device=self.device)
                                   for k, dt, shp, off, per_img in self.layout}
            full = self.full[slot]
            ready = torch.cuda.Event()
            ready.record(cur)
            streams = self.side
            for s in streams:
                s.wait_event(ready)
                s.wait_event(self.done[slot])              # previous consumer of this slot's full-batch buffers (none the first time)
            o, n = self.offsets[self.rank], self.counts[self.rank]
            with torch.cuda.stream(streams[0]):
                for k, dt, shp, off, per_img in self.layout:
                    t = local[k]
                    full[k][o:o + n].copy_(t.view(torch.uint8) if t.dtype == torch.bool else t, non_blocking=True)
            for t in local.values():
                t.record_stream(streams[0])
            for i, (r, base) in enumerate(sorted(self.peers.items())):
                s = streams[1 + i % (len(streams) - 1)] if len(streams) > 1 else streams[0]
                st = s.cuda_stream
                tokw = self._mark("dst: wait for a peer's flag", s)
                capi.check(L.moge_peer_flag_wait(self._ready_flag(base, slot), seq, st))
                self._end(tokw, s)
                tok = self._mark("dst: pull one peer", s)
                for k, dt, shp, off, per_img in self.layout:
                    dstp = full[k].data_ptr() + self.offsets[r] * per_img
                    capi.check(L.moge_peer_copy(dstp, base + slot * self.slot_bytes + off, per_img * self.counts[r], st))
                self._end(tok, s)
            # all pulls of this slot done -> producers may reuse it
            fin = streams[0]
            for s in streams[1:]:
                ev = torch.cuda.Event()
                ev.record(s)
                fin.wait_event(ev)
            capi.check(L.moge_peer_flag_set(self.flag_base + 4 * (self.SLOTS + slot), seq, fin.cuda_stream))
            self.done[slot].record(fin)
            return {k: (v.view(torch.bool) if dt == torch.bool else v) for (k, dt, shp, off, per_img), v in
                    ((e, full[e[0]]) for e in self.layout)}

    def fence(self, stream=None) -> None:
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream()
            for ev in self.done:
                st.wait_event(ev)

    def wait(self) -> None:
        self.fence()
        torch.cuda.current_stream(self.device).synchronize()

    def close(self) -> None:
        L = self.L
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self.peers.values():
                L.moge_peer_close(p)
            if self.flags_dst is not None:
                L.moge_peer_close(self.flags_dst - self.SLOTS * self.slot_bytes)
            if self.stage:
                L.moge_peer_free(self.stage)
        self.peers, self.flags_dst, self.stage = {}, None, None
