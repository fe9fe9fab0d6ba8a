"""Throughput path for HOST-resident batches: the host->device copy of batch i+1, `infer()` of batch i and the device->host
copy of the outputs of batch i-1 run on three CUDA streams (double-buffered), so PCIe time disappears behind the network.

The reference has no counterpart (its scripts call `model.infer` on one image at a time, moge/scripts/infer.py:101); this is
the serving loop a deployment puts around `MoGeModel.infer`, and what `bench.py` reports as `e2e`.

    pipe = InferPipeline(model, num_tokens=1369)
    for x, out in zip(pinned_inputs, pinned_output_dicts):
        pipe.submit(x, out)          # returns immediately
    pipe.join()                      # all outputs are in the pinned host buffers
"""
from typing import Dict, List, Optional

import torch


class InferPipeline:
    def __init__(self, model, depth: int = 2, **infer_kwargs):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model = model
        self.kw = infer_kwargs
        self.depth = depth
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("InferPipeline needs the model on a CUDA device (there is no CPU path)")
        self.dev = dev
        with torch.cuda.device(dev):
            self.h2d = torch.cuda.Stream()
            self.compute = torch.cuda.Stream()
            self.d2h = torch.cuda.Stream()
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]
            self.ev_compute = [torch.cuda.Event() for _ in range(depth)]
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.dev_in: List[Optional[torch.Tensor]] = [None] * depth
        self.n = 0

    def submit(self, host_in: torch.Tensor, host_out: Dict[str, torch.Tensor]) -> None:
        """Queue one batch.  `host_in` (B,3,H,W) and the tensors of `host_out` (keys of `infer()`'s result) should be pinned;
        `host_out` must not be read before `join()` (or a synchronize of `self.d2h`)."""
        s = self.n % self.depth
        self.n += 1
        with torch.cuda.device(self.dev):
            self.h2d.wait_event(self.ev_compute[s])            # the previous user of this input slot has been consumed
            with torch.cuda.stream(self.h2d):
                if self.dev_in[s] is None or self.dev_in[s].shape != host_in.shape or self.dev_in[s].dtype != host_in.dtype:
                    # (re)allocated ON the copy stream, after the wait above: the old buffer goes back to that stream's pool only
                    # once its last reader (the compute stream) is done, and the new one is known to the compute stream below
                    self.dev_in[s] = torch.empty(host_in.shape, dtype=host_in.dtype, device=self.dev)
                    self.dev_in[s].record_stream(self.compute)
                self.dev_in[s].copy_(host_in, non_blocking=True)
                self.ev_in[s].record(self.h2d)
            self.compute.wait_event(self.ev_in[s])
            with torch.cuda.stream(self.compute):
                out = self.model.infer(self.dev_in[s], **self.kw)
                self.ev_compute[s].record(self.compute)
            self.d2h.wait_event(self.ev_compute[s])
            with torch.cuda.stream(self.d2h):
                for k, v in out.items():
                    host_out[k].copy_(v, non_blocking=True)
                    v.record_stream(self.d2h)                  # allocated on the compute stream, last read here
                self.ev_out[s].record(self.d2h)

    def join(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Make `stream` (default: the current stream) wait for everything submitted so far, then block the host on it."""
        with torch.cuda.device(self.dev):
            st = stream if stream is not None else torch.cuda.current_stream()
            for s in range(self.depth):
                st.wait_event(self.ev_out[s])
            st.synchronize()

    def fence(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Device-side only: `stream` waits for everything submitted so far (no host block)."""
        with torch.cuda.device(self.dev):
            st = stream if stream is not None else torch.cuda.current_stream()
            for s in range(self.depth):
                st.wait_event(self.ev_out[s])
