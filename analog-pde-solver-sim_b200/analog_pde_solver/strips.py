"""Row-strip decomposition of the red-black throughput path across GPUs (one process per GPU).

The reference is single-process; this is the multi-GPU driver of the new backend.
Rank r owns a contiguous block of global rows plus ``2*temporal_depth`` ghost rows on each
interior side.  A *group* of up to ``temporal_depth`` sweeps runs between halo exchanges: after the
group each rank sends its outermost ``halo`` owned rows to the neighbour's ghost rows
(``torch.distributed`` P2P: NCCL over NVLink on GPUs, gloo in the CPU tests) and the per-sweep
residual maxima are all-reduced (MAX).  The arithmetic per node is identical to the single-strip
run, so the multi-GPU result equals the 1-GPU result bit for bit.

The sweep engine is injected: ``CudaStripEngine`` wraps ``libapde.so`` plans; the CPU tests pass a
stand-in built on the test oracle to exercise the host logic without a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def partition_rows(grows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row ranges [begin, end) for ranks 0..world-1."""
    base, extra = divmod(grows, world)
    out, begin = [], 0
    for r in range(world):
        end = begin + base + (1 if r < extra else 0)
        out.append((begin, end))
        begin = end
    return out


def overlapped_strip(begin: int, end: int, grows: int, n_sweeps: int) -> Tuple[int, int, int, int]:
    """Communication-avoiding strip for a FIXED number of sweeps: the rows [lo, hi) a rank relaxes on its own and the
    window [win_lo, win_hi) inside them (local indices) that reproduces the whole grid's sweeps bit for bit.

    A red-black sweep moves information 2 rows, so after n_sweeps sweeps only rows within 2*n_sweeps - 1 of the artificial
    (frozen) first / last row of the extended strip differ from the global solution.  The extension G is the smallest even
    number >= 2*n_sweeps (even: the strip's own colouring (i + j) % 2 must be the grid's), clipped at the grid's real
    boundary.  Feed `lo:hi` of the host arrays to `_backend.solve_window(..., win_lo, win_hi)` and max-combine the residual
    histories of the ranks (include/apde.h: apde_solve_redblack_window)."""
    if not (0 <= begin < end <= grows) or n_sweeps < 0:
        raise ValueError("overlapped_strip: need 0 <= begin < end <= grows and n_sweeps >= 0")
    g = 2 * n_sweeps
    g += g % 2
    lo = max(begin - g, 0)
    if lo % 2:  # keep the colouring: an odd start moves one more row out (or in, at the boundary it is row 0 anyway)
        lo -= 1
    hi = min(end + g, grows)
    return lo, hi, begin - lo, end - lo


def sweep_groups(n_sweeps: int, depth: int) -> List[int]:
    """Split n_sweeps into groups of at most `depth` sweeps (one halo exchange per group)."""
    full, rest = divmod(n_sweeps, depth)
    return [depth] * full + ([rest] if rest else [])


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap a raw device pointer."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


class CudaStripEngine:
    """One libapde plan (device-resident strip) + torch views of its halo rows and residuals."""

    def __init__(self, grows, cols, row_begin, row_end, dtype=np.float64, has_noise=False,
                 has_source=False, temporal_depth=4):
        import torch
        from . import _backend
        self.torch = torch
        self.plan = _backend.Plan(grows, cols, row_begin, row_end, dtype=dtype, has_noise=has_noise,
                                  has_source=has_source, temporal_depth=temporal_depth)
        self.dtype = np.dtype(dtype)

    def stream(self) -> int:
        return int(self.torch.cuda.current_stream().cuda_stream)

    def run(self, n_sweeps, sweep_base, phase=0):
        self.plan.run(n_sweeps, sweep_base, phase, self.stream())

    def clear_residuals(self, first, count):
        self.plan.clear_residuals(first, count, self.stream())

    def halo_tensors(self, side):
        send, recv, nbytes = self.plan.halo(side)
        if send is None:
            return None, None
        wrap = lambda p: self.torch.as_tensor(_CudaArray(p, nbytes), device="cuda")  # noqa: E731
        return wrap(send), wrap(recv)

    def residual_tensor(self, first, count):
        import ctypes
        from . import _backend
        ptr, cap = ctypes.c_void_p(None), ctypes.c_int64(0)
        _backend.check(self.plan._lib.apde_plan_residuals(self.plan._h, ctypes.byref(ptr), ctypes.byref(cap)))
        item = self.dtype.itemsize
        raw = self.torch.as_tensor(_CudaArray(ptr.value + first * item, count * item), device="cuda")
        return raw.view(self.torch.float64 if item == 8 else self.torch.float32)

    def comm_scope(self):
        """Side stream for the halo exchange: it first waits for the work queued so far on the
        main stream (the strip-edge row blocks), then carries the NCCL send/recv while the main
        stream runs the interior row blocks."""
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = self.torch.cuda.Stream(priority=-1)  # edge blocks + halo traffic go first
        self._comm_stream.wait_stream(self.torch.cuda.current_stream())
        return self.torch.cuda.stream(self._comm_stream)

    def join_comm(self):
        if hasattr(self, "_comm_stream"):
            self.torch.cuda.current_stream().wait_stream(self._comm_stream)

    def close(self):
        self.plan.close()


class StripSolver:
    """Drives one strip per rank.  `engine` must offer run / halo_tensors / residual_tensor."""

    def __init__(self, engine, rank: int, world: int, temporal_depth: int, dist=None):
        self.engine = engine
        self.rank = rank
        self.world = world
        self.depth = temporal_depth
        self.dist = dist  # torch.distributed module (initialised) or None for a single strip
        self.sweeps_done = 0
        self.exchanges = 0

    def exchange_halos(self):
        """Send my outermost owned rows up/down, receive the neighbours' into my ghost rows."""
        if self.dist is None or self.world == 1:
            return
        ops = []
        for side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
            if peer < 0 or peer >= self.world:
                continue
            send, recv = self.engine.halo_tensors(side)
            ops.append(self.dist.P2POp(self.dist.isend, send, peer))
            ops.append(self.dist.P2POp(self.dist.irecv, recv, peer))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        self.exchanges += 1

    def run(self, n_sweeps: int, overlap: bool = True):
        """Run n_sweeps sweeps (groups of <= temporal_depth, one exchange after each group)."""
        for n in sweep_groups(n_sweeps, self.depth):
            if overlap and self.world > 1:
                # fork: the high-priority side stream relaxes the strip-edge row blocks and ships
                # them to the neighbours while the main stream relaxes the interior row blocks;
                # both read the same input buffer and write disjoint rows of the output buffer
                self.engine.clear_residuals(self.sweeps_done, n)
                with self.engine.comm_scope():
                    self.engine.run(n, self.sweeps_done, phase=1)
                    self.exchange_halos()
                self.engine.run(n, self.sweeps_done, phase=2)
                self.engine.join_comm()
            else:
                self.engine.run(n, self.sweeps_done, phase=0)
                self.exchange_halos()
            self.sweeps_done += n

    def analog_energy_joules(self, resistor_ohms: float = 10_000.0) -> float:
        """EnergyModel.analog_energy_joules (upstream solver.py:70-97) without moving the grid off the
        GPUs: per-strip sums of squared link drops, SUM all-reduce, same closing arithmetic."""
        sum_h, sum_v = self.engine.plan.energy_sums(self.engine.stream())
        if self.dist is not None and self.world > 1:
            t = self.engine.torch.tensor([sum_h, sum_v], dtype=self.engine.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            sum_h, sum_v = (float(x) for x in t.tolist())
        watts = sum_h / resistor_ohms + sum_v / resistor_ohms
        return watts * (5.0 * resistor_ohms * 1e-12)

    def residuals(self, first: int, count: int) -> np.ndarray:
        """Global per-sweep max_change: MAX all-reduce of the ranks' device arrays."""
        t = self.engine.residual_tensor(first, count)
        if self.dist is not None and self.world > 1:
            t = t.clone()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.detach().cpu().numpy().astype(np.float64)
