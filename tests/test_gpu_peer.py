"""-m gpu, needs two GPUs: the peer-memory output gather (moge_b200.parallel.PeerGatherer over the moge_peer_* C ABI): CUDA-IPC
staging slots, copy-engine pulls, device-side flags.  Two processes (one per GPU); gloo carries only the IPC handles."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


class _FailingAlloc:
    """the C library with one entry point failing (fault injection for the collective error path)"""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def moge_peer_alloc(self, *args):
        return -1


def _worker(rank, world, port, q, fail_rank=-1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    try:
        from moge_b200.parallel import PeerGatherer, PeerSetupError, OutputGatherer, shard_range
        if fail_rank >= 0:
            # one rank cannot allocate its staging memory: EVERY rank must raise PeerSetupError at the first submit (nobody is
            # left polling a device flag), and the portable gatherer must work right afterwards on the same process group
            dev = torch.device("cuda", rank)
            torch.cuda.set_device(dev)
            gat = PeerGatherer([2, 2], dev)
            if rank == fail_rank:
                gat.L = _FailingAlloc(gat.L)
            local = {"points": torch.full((2, 5, 7, 3), float(rank + 1), device=dev)}
            try:
                gat.submit(local)
                ok = False
            except PeerSetupError as ex:
                ok = ok and f"rank {fail_rank}" in str(ex)
            fb = OutputGatherer([2, 2])                 # gloo group of this test: host tensors
            res = fb.submit({k: v.cpu() for k, v in local.items()})
            fb.wait()
            if rank == 0:
                ok = ok and torch.equal(res["points"][:2].cpu(), torch.full((2, 5, 7, 3), 1.0))
                ok = ok and torch.equal(res["points"][2:].cpu(), torch.full((2, 5, 7, 3), 2.0))
            dist.barrier()
            return
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        total = 5
        counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
        lo, hi = shard_range(total, rank, world)
        g = torch.Generator().manual_seed(3)
        pts = torch.randn(total, 37, 41, 3, generator=g)
        msk = torch.rand(total, 37, 41, generator=g) > 0.4
        K = torch.randn(total, 3, 3, generator=g)
        gat = PeerGatherer(counts, dev)
        for step in range(7):                      # more steps than slots: exercises the "consumed" back-pressure
            local = {"points": (pts[lo:hi] + step).to(dev), "mask": (msk[lo:hi] if step % 2 == 0 else ~msk[lo:hi]).to(dev),
                     "intrinsics": (K[lo:hi] * (step + 1)).to(dev)}
            res = gat.submit(local)
            if step in (2, 6):
                gat.wait()
                if rank == 0:
                    ok = ok and torch.equal(res["points"].cpu(), pts + step) and res["mask"].dtype == torch.bool
                    ok = ok and torch.equal(res["mask"].cpu(), msk if step % 2 == 0 else ~msk)
                    ok = ok and torch.equal(res["intrinsics"].cpu(), K * (step + 1))
                else:
                    ok = ok and res is None
        gat.wait()
        dist.barrier()
        gat.close()
    except Exception as e:      # noqa: BLE001
        import traceback
        traceback.print_exc()
        ok = False
    finally:
        q.put((rank, bool(ok)))
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_gather_two_gpus():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_setup_failure_is_raised_on_every_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, 1)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
