"""Collectives for world > 1: the two host-array all-reduces the C ABI asks for (scs_set_collectives), on top of
torch.distributed. One process per GPU; backend NCCL on the GPU box (arrays staged through a device tensor, NVLink),
gloo on CPU for the host-logic tests. The path has no bulk exchange: per amplification round a 2-word and a 1-word sum,
per batch one word per rank, once the weight vector of the cell."""
from __future__ import annotations

import numpy as np


def make_collectives(dist, device="cuda"):
    """Return (allreduce_u64, allreduce_f64): in-place sums over all ranks of a 1-D numpy array."""
    import torch

    def _sum(a: np.ndarray, torch_dtype):
        t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a)
        if device != "cpu":
            d = t.to(device)
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
            t.copy_(d.cpu())
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def ar_u64(a: np.ndarray):   # two's-complement add == unsigned add
        _sum(a, torch.int64)

    def ar_f64(a: np.ndarray):
        _sum(a, torch.float64)

    return ar_u64, ar_f64


def make_device_allreduce(dist, device):
    """Return (f64, i64): f(ptr, n) = in-place NCCL sum of n 8-byte elements in device memory at `ptr` (zero-copy view
    through __cuda_array_interface__), finished when it returns."""
    import torch

    class _View:
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

    def make(typestr):
        def ar(ptr: int, n: int):
            t = torch.as_tensor(_View(ptr, n, typestr), device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            torch.cuda.current_stream(t.device).synchronize()
        return ar

    return make("<f8"), make("<i8")


class ThreadCollectives:
    """In-process stand-in (ranks = threads sharing one GPU) used by the single-GPU multi-rank parity test."""

    def __init__(self, world: int):
        import threading
        self.world, self.lock, self.bar = world, threading.Lock(), threading.Barrier(world)
        self.acc = None

    def _sum(self, a: np.ndarray):
        self.bar.wait()
        with self.lock:
            if self.acc is None:
                self.acc = a.copy()
            else:
                self.acc += a
        self.bar.wait()
        a[:] = self.acc
        self.bar.wait()
        with self.lock:
            self.acc = None

    def pair(self):
        return self._sum, self._sum

    def device_pair(self, device="cuda:0"):
        """Device-memory hooks for the in-process test: staged through the host sums above."""
        import torch

        class _View:
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

        def make(typestr):
            def ar(ptr, n):
                t = torch.as_tensor(_View(ptr, n, typestr), device=device)
                h = t.cpu().numpy().copy()
                self._sum(h)
                t.copy_(torch.from_numpy(h))
                torch.cuda.synchronize()
            return ar
        return make("<f8"), make("<i8")
