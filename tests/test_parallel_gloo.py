"""CPU, world_size=2, gloo: the data-parallel host logic (shard / broadcast weights / gather outputs)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moge_b200.parallel import shard_range, broadcast_state_dict, gather_outputs, OutputGatherer


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 32, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        state = {"a.weight": torch.randn(5, 3, generator=g), "b.bias": torch.randn(7, generator=g), "n": torch.tensor(3)} if rank == 0 else None
        got = broadcast_state_dict(state, torch.device("cpu"))
        ref = {"a.weight": torch.randn(5, 3, generator=torch.Generator().manual_seed(0))}
        ok = torch.equal(got["a.weight"], ref["a.weight"]) and set(got) == {"a.weight", "b.bias"}
        total = 5
        counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
        lo, hi = shard_range(total, rank, world)
        full_pts = torch.arange(total * 2 * 3, dtype=torch.float32).view(total, 2, 3)
        full_mask = (torch.arange(total * 2) % 3 == 0).view(total, 2)
        out = gather_outputs({"points": full_pts[lo:hi].clone(), "mask": full_mask[lo:hi].clone()}, counts)
        if rank == 0:
            ok = ok and torch.equal(out["points"], full_pts) and torch.equal(out["mask"], full_mask) and out["mask"].dtype == torch.bool
        else:
            ok = ok and out is None
        # pipelined gatherer (grouped isend/irecv into preallocated full-batch buffers), several steps through 2 slots
        gat = OutputGatherer(counts)
        for step in range(3):
            res = gat.submit({"points": full_pts[lo:hi] + step, "mask": (full_mask[lo:hi] if step % 2 == 0 else ~full_mask[lo:hi]).clone()})
            gat.wait()
            if rank == 0:
                ok = ok and torch.equal(res["points"], full_pts + step) and res["mask"].dtype == torch.bool
                ok = ok and torch.equal(res["mask"], full_mask if step % 2 == 0 else ~full_mask)
            else:
                ok = ok and res is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
