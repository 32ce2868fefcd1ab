#!/usr/bin/env python3
"""CPU model of the streamed one-shot solve's schedule (csrc/apde_redblack.cu, solve_streamed_t): the same band edges,
upload chunks, lanes and events, restated in Python, and two checks that do not depend on timing:

  1. happens-before: for every pair of operations that the streams and events leave UNORDERED, their accesses to the two
     ping-pong grids and to the operand planes do not conflict (no write/write, no read/write overlap of rows);
  2. values: executed in any order the schedule allows, every pass of every band reads, on all rows it touches
     (band + 2*TD halo rows either side), exactly the previous pass's values (pass 0 = the uploaded grid), and every
     downloaded row holds the last pass.

Together: the streamed result equals the unstreamed one for every interleaving the GPU may pick.  The GPU tests check the
same thing bit for bit on real runs (tests/test_redblack_gpu.py::test_streamed_*); this model covers shapes they do not
(shifts beyond a band, bands that vanish, one to two lanes, tiny last bands).  Usage: python tools/stream_schedule_sim.py"""
import random
import sys


def band_edges(rows, B, h, P):
    """edge0 of solve_streamed_t: B/4, B/2, B, ..., B, B/2, B/4 until rows + (P-1)*h is covered."""
    skew_max = (P - 1) * h
    q = max(4 * h, (B // 4) & ~7)
    hq = max(4 * h, (B // 2) & ~7)
    need = rows + skew_max
    n_full = max(0, -(-(need - 2 * (q + hq)) // B))
    sizes = [q, hq] + [B] * n_full + [hq, q]
    edge0 = [0]
    for sz in sizes:
        if edge0[-1] >= need:
            break
        edge0.append(edge0[-1] + sz)
    return edge0


def build(rows, B, td, n_sweeps, lanes=2, mutate=None):
    """mutate (negative controls): 'no_lane_event' drops the wait on the other lane's previous pass, 'short_shift' shifts the
    bands by one row less than the halo, 'early_band' lets a band start one upload chunk early."""
    h = 2 * td
    P = -(-n_sweeps // td)
    edge0 = band_edges(rows, B, h, P)
    nb = len(edge0) - 1
    nchunk = nb
    clamp = lambda x: min(rows, max(0, x))
    shift = h - 1 if mutate == "short_shift" else h
    band_edge = lambda b, p: clamp(edge0[min(b, nb)] - (p - 1) * shift)
    chunk_edge = lambda c: 0 if c <= 0 else min(rows, edge0[min(c, nb)] + h)
    ops = []      # dict(stream, reads=[(buf, lo, hi)], writes=[...], waits=[op index], kind, ...)
    last_in_stream = {}

    def add(stream, kind, reads, writes, waits=(), **kw):
        i = len(ops)
        w = list(waits)
        if stream in last_in_stream:
            w.append(last_in_stream[stream])
        ops.append(dict(stream=stream, kind=kind, reads=reads, writes=writes, waits=w, **kw))
        last_in_stream[stream] = i
        return i

    up_done, band_done = [None] * nchunk, [None] * nb
    pass_done = [[None] * (P + 1), [None] * (P + 1)]
    for i in range(max(nchunk, nb) + 1):
        if i < nchunk:
            r0, r1 = chunk_edge(i), chunk_edge(i + 1)
            wr = [("U0", r0, r1), ("U1", r0, r1), ("S", r0, r1)] if r1 > r0 else []
            up_done[i] = add("up", "upload", [], wr, chunk=i)
        if i < nb:
            b = i
            st = "lane1" if (lanes == 2 and b % 2) else "lane0"
            prev, mine = pass_done[(b + 1) % 2], pass_done[b % 2]
            first = True
            for p in range(1, P + 1):
                lo = band_edge(b, p)
                hi = rows if b == nb - 1 else band_edge(b + 1, p)
                waits = []
                if first:
                    waits.append(up_done[max(0, min(b, nchunk - 1) - (1 if mutate == "early_band" else 0))])
                    first = False
                if lanes == 2 and b > 0 and p > 1 and prev[p - 1] is not None and mutate != "no_lane_event":
                    waits.append(prev[p - 1])
                src, dst = "U%d" % ((p - 1) % 2), "U%d" % (p % 2)
                if hi > lo:
                    rlo, rhi = max(0, lo - h), min(rows, hi + h)
                    k = add(st, "pass", [(src, rlo, rhi), ("S", rlo, rhi)], [(dst, lo, hi)], waits, band=b, p=p, lo=lo, hi=hi)
                else:
                    k = add(st, "nop", [], [], waits, band=b, p=p)
                mine[p] = k  # the event recorded behind this pass
            band_done[b] = last_in_stream[st]
        if i >= 1 and i - 1 < nb:
            b = i - 1
            lo = band_edge(b, P)
            hi = rows if b == nb - 1 else band_edge(b + 1, P)
            if hi > lo:
                add("down", "download", [("U%d" % (P % 2), lo, hi)], [], [band_done[b]], band=b, lo=lo, hi=hi)
    return ops, P, nb


def check(rows, B, td, n_sweeps, lanes=2, seed=0, mutate=None):
    ops, P, nb = build(rows, B, td, n_sweeps, lanes, mutate)
    n = len(ops)
    # happens-before closure (ops are created in an order compatible with their waits)
    before = [0] * n
    for i, o in enumerate(ops):
        m = 0
        for w in o["waits"]:
            assert w < i
            m |= before[w] | (1 << w)
        before[i] = m
    # 1. unordered pairs must not conflict
    acc = []
    for o in ops:
        acc.append(([(b, lo, hi, False) for b, lo, hi in o["reads"]] + [(b, lo, hi, True) for b, lo, hi in o["writes"]]))
    for i in range(n):
        for j in range(i):
            if (before[i] >> j) & 1:
                continue
            for (b1, l1, h1, w1) in acc[i]:
                for (b2, l2, h2, w2) in acc[j]:
                    if b1 == b2 and (w1 or w2) and l1 < h2 and l2 < h1:
                        raise AssertionError(f"unordered conflict on {b1} rows [{max(l1, l2)},{min(h1, h2)}): {ops[j]} vs {ops[i]}")
    # 2. values, in a random order the waits allow
    rng = random.Random(seed)
    ver = {"U0": [None] * rows, "U1": [None] * rows, "S": [None] * rows}
    done, pending, got = set(), list(range(n)), [None] * rows
    while pending:
        ready = [i for i in pending if all(w in done for w in ops[i]["waits"])]
        i = rng.choice(ready)
        o = ops[i]
        if o["kind"] == "upload":
            for b, lo, hi in o["writes"]:
                for r in range(lo, hi):
                    ver[b][r] = 0
        elif o["kind"] == "pass":
            p = o["p"]
            for b, lo, hi in o["reads"]:
                want = 0 if b == "S" else p - 1
                for r in range(lo, hi):
                    if ver[b][r] != want:
                        raise AssertionError(f"band {o['band']} pass {p} reads {b} row {r}: version {ver[b][r]}, wants {want}")
            for b, lo, hi in o["writes"]:
                for r in range(lo, hi):
                    ver[b][r] = p
        elif o["kind"] == "download":
            b, lo, hi = o["reads"][0]
            for r in range(lo, hi):
                assert ver[b][r] == P, f"download of row {r}: version {ver[b][r]} instead of {P}"
                assert got[r] is None
                got[r] = P
        done.add(i)
        pending.remove(i)
    assert all(g == P for g in got), "rows never downloaded"
    return n


if __name__ == "__main__":
    shapes = [(700, 64, 4, 37), (515, 128, 3, 24), (2100, 256, 4, 130), (16384, 1536, 4, 256), (300, 96, 4, 9), (1300, 64, 4, 20)]
    for rows, B, td, k in shapes:
        for lanes in (1, 2):
            if rows * k > 3_000_000 and lanes == 1:
                continue
            n = check(rows, B, td, k, lanes)
            print(f"rows {rows} band {B} depth {td} sweeps {k} lanes {lanes}: {n} operations, no unordered conflict, every read sees the previous pass", flush=True)
    sys.exit(0)
