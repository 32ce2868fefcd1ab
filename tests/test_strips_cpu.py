"""CPU (gloo, world_size 2 and 3): host logic of the row-strip driver -- partitioning, sweep groups,
halo exchange order, residual MAX all-reduce.  The sweep engine is a stand-in built on the test
oracle (the product engine is CudaStripEngine -> libapde.so; tests/test_redblack_gpu.py and
tests/test_strips_gpu.py cover it on the GPU).  The multi-rank result must equal the single-strip
oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import random_problem
from analog_pde_solver.strips import StripSolver, partition_rows, sweep_groups


class OracleStripEngine:
    """Local strip with ghost rows, relaxed by the CPU oracle (test stand-in for a libapde plan)."""

    def __init__(self, g0, nh, nv, src, row_begin, row_end, depth, dtype=np.float64):
        grows, cols = g0.shape
        self.halo = 2 * depth
        self.gt = self.halo if row_begin > 0 else 0
        self.gb = self.halo if row_end < grows else 0
        lo, hi = row_begin - self.gt, row_end + self.gb
        self.pad = lo & 1  # keep the oracle's local colour parity equal to the global one
        sl = slice(lo - self.pad, hi)
        self.u = np.ascontiguousarray(g0[sl]).astype(dtype)
        self.nh = None if nh is None else np.ascontiguousarray(nh[sl])
        self.nv = None if nv is None else np.ascontiguousarray(nv[lo - self.pad:hi - 1])
        self.src = None if src is None else np.ascontiguousarray(src[sl])
        self.own = slice(self.pad + self.gt, self.pad + self.gt + (row_end - row_begin))
        self.dtype = dtype
        self.resid = np.zeros(4096, dtype=np.float64)
        self.calls = []

    def run(self, n, base, phase=0):
        self.calls.append((n, base, phase))
        if phase == 2:
            return
        for s in range(n):
            before = self.u.copy()
            self.u, _, _ = oracle.red_black(self.u, self.nh, self.nv, self.src, 1, dtype=self.dtype)
            ch = np.abs(self.u[self.own].astype(np.float64) - before[self.own].astype(np.float64))
            ch = ch[~np.isnan(ch)]
            self.resid[base + s] = ch.max() if ch.size else 0.0

    def clear_residuals(self, first, count):
        self.resid[first:first + count] = 0.0

    def halo_tensors(self, side):
        t = torch.from_numpy(self.u)
        if side == 0 and self.gt:
            a = self.pad
            return t[a + self.gt:a + self.gt + self.halo], t[a:a + self.gt]
        if side == 1 and self.gb:
            e = self.own.stop
            return t[e - self.halo:e], t[e:e + self.gb]
        return None, None

    def residual_tensor(self, first, count):
        return torch.from_numpy(self.resid[first:first + count])

    # the CPU stand-in has no streams
    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def comm_scope(self):
        return self._Null()

    def join_comm(self):
        pass


def _worker(rank, world, port, rows, cols, depth, n_sweeps, overlap, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g0, nh, nv, src = random_problem(rows, cols, 0.01, seed=5)
        begin, end = partition_rows(rows, world)[rank]
        eng = OracleStripEngine(g0, nh, nv, src, begin, end, depth)
        solver = StripSolver(eng, rank, world, depth, dist)
        solver.run(n_sweeps, overlap=overlap)
        res = solver.residuals(0, n_sweeps)
        ref, ref_hist, _ = oracle.red_black(g0, nh, nv, src, n_sweeps)
        ok_grid = eng.u[eng.own].tobytes() == ref[begin:end].tobytes()
        ok_hist = res.tobytes() == ref_hist.tobytes()
        out[rank] = (ok_grid, ok_hist, solver.exchanges, [c[2] for c in eng.calls])
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,rows,cols,depth,n_sweeps,overlap", [
    (2, 40, 22, 2, 7, False), (2, 41, 19, 3, 6, True), (3, 60, 17, 2, 5, True)])
def test_strips_equal_single_strip_bitwise(world, rows, cols, depth, n_sweeps, overlap):
    mgr = mp.get_context("spawn").Manager()  # never fork a process that already runs OpenMP threads
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), rows, cols, depth, n_sweeps, overlap, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok_grid, ok_hist, exchanges, phases = out[rank]
        assert ok_grid, f"rank {rank}: owned rows differ from the single-strip result"
        assert ok_hist, f"rank {rank}: all-reduced residuals differ"
        assert exchanges == len(sweep_groups(n_sweeps, depth))
        assert set(phases) == ({1, 2} if overlap else {0})


def test_partition_and_groups():
    assert partition_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert partition_rows(16384 * 8, 8)[7] == (16384 * 7, 16384 * 8)
    parts = partition_rows(65536, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 65536 and all(b - a == 8192 for a, b in parts)
    assert sweep_groups(32, 3) == [3] * 10 + [2]
    assert sweep_groups(4, 4) == [4] and sweep_groups(0, 4) == []


def test_overlapped_strips_reproduce_the_whole_grid_without_halo_exchange():
    """The dependence-cone statement behind apde_solve_redblack_window, on the CPU oracle: every strip, extended by
    2*K ghost rows per interior side and relaxed K sweeps ON ITS OWN (first / last row frozen), equals the whole grid's
    result inside its window bit for bit; the maximum of the strips' windowed residuals is the whole grid's residual."""
    import oracle
    from helpers import random_problem
    from analog_pde_solver.strips import overlapped_strip, partition_rows
    rows, cols, k = 181, 37, 9
    g0, nh, nv, s = random_problem(rows, cols, 0.02, seed=4, with_source=True)
    ref, ref_hist, _ = oracle.red_black(g0, nh, nv, s, k)
    out = g0.copy()
    for world in (2, 3, 5):
        for begin, end in partition_rows(rows, world):
            lo, hi, wl, wh = overlapped_strip(begin, end, rows, k)
            assert lo % 2 == 0 and (lo == 0 or begin - lo >= 2 * k) and (hi == rows or hi - end >= 2 * k)
            sub, _, _ = oracle.red_black(g0[lo:hi].copy(), nh[lo:hi].copy(), nv[lo:hi - 1].copy(), s[lo:hi].copy(), k)
            out[begin:end] = sub[wl:wh]
        assert out.tobytes() == ref.tobytes(), world
    # one ghost row too few is not enough: the cone is exactly 2K - 1 rows deep
    begin, end = partition_rows(rows, 2)[1]
    lo = begin - (2 * k - 2)
    sub, _, _ = oracle.red_black(g0[lo:].copy(), nh[lo:].copy(), nv[lo:].copy(), s[lo:].copy(), k)
    assert sub[begin - lo:].tobytes() != ref[begin:].tobytes()


def test_streamed_schedule_model_has_no_unordered_conflicts_and_reads_the_right_pass():
    """tools/stream_schedule_sim.py: the band edges / chunks / lanes / events of the streamed one-shot solve, restated on the
    CPU.  For random shapes every pair of operations the streams leave unordered touches disjoint rows, and every pass
    reads exactly the previous pass on band + halo; dropping the cross-lane event, shifting the bands by less than the
    halo, or starting a band one upload chunk early is caught."""
    import importlib.util, os, random
    spec = importlib.util.spec_from_file_location("stream_schedule_sim", os.path.join(os.path.dirname(__file__), "..", "tools", "stream_schedule_sim.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    rng = random.Random(7)
    for _ in range(40):
        td = rng.choice([1, 2, 3, 4])
        B = rng.choice([32, 64, 96, 128, 256]) 
        if B < 8 * td:
            continue
        rows = rng.randint(2 * B, 12 * B)
        k = rng.randint(1, 90)
        sim.check(rows, B, td, k, lanes=rng.choice([1, 2]), seed=rng.randint(0, 99))
    sim.check(16384, 1536, 4, 256, lanes=2)  # the benchmark shape
    for mutate in ("no_lane_event", "short_shift", "early_band"):
        with pytest.raises(AssertionError):
            sim.check(900, 64, 4, 40, lanes=2, mutate=mutate)


def _ca_worker(rank, world, port, rows, cols, k, out):
    """One rank of the communication-avoiding driver (what bench.py does at N > 1 for its end-to-end figure): relax the
    extended strip on its own -- here with the oracle -- take the residual maxima over the owned rows only, MAX all-reduce."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from analog_pde_solver.strips import overlapped_strip
        g0, nh, nv, src = random_problem(rows, cols, 0.01, seed=11)
        begin, end = partition_rows(rows, world)[rank]
        lo, hi, wl, wh = overlapped_strip(begin, end, rows, k)
        u = g0[lo:hi].copy()
        hist = np.zeros(k)
        for s in range(k):  # sweep by sweep, to take the windowed residual like apde_solve_redblack_window
            before = u
            u, _, _ = oracle.red_black(u, nh[lo:hi].copy(), nv[lo:hi - 1].copy(), src[lo:hi].copy(), 1)
            hist[s] = np.abs(u[wl:wh] - before[wl:wh]).max()
        t = torch.from_numpy(hist)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ref, ref_hist, _ = oracle.red_black(g0, nh, nv, src, k)
        out[rank] = (u[wl:wh].tobytes() == ref[begin:end].tobytes(), t.numpy().tobytes() == ref_hist.tobytes(), hi - lo)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rows,cols,k", [(2, 90, 21, 7), (3, 151, 18, 6)])
def test_communication_avoiding_strips_gloo(world, rows, cols, k):
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_ca_worker, args=(world, _free_port(), rows, cols, k, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok_grid, ok_hist, sub_rows = out[rank]
        assert ok_grid, f"rank {rank}: window rows differ from the whole-grid result"
        assert ok_hist, f"rank {rank}: max-combined windowed residuals differ from the whole grid's"
