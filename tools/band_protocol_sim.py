#!/usr/bin/env python3
"""CPU replay of the progress-flag protocol of the exact path's row-band kernel (K2c, csrc/apde_exact.cu
`exact_band_kernel`): every CTA (sweep slot, band) publishes the steps it has completed every `kb` steps and, before a
batch of `kb` steps, waits for three predecessors.  The replay advances any CTA that may advance until nothing moves;
CTAs left short of their last step mean the launch shape would deadlock on the GPU.  Used by tests/test_abi.py to pin
the host's bound on the number of sweep slots (slots * (3*kb + 4) <= cols) -- a cooperative kernel that deadlocks
hangs the GPU, so the bound is checked where no GPU is needed."""
import random
import sys


def launch_shape(rows, cols, nt, kb, k_sweeps, grid_max):
    """(bands, slots) as ring_configure / solve_ring choose them; nt = band height (threads x rows per thread)."""
    bands = (rows - 2 + nt - 1) // nt
    slots_max = min(grid_max // bands, cols // (3 * kb + 4))
    if slots_max < 1:
        return bands, 0
    rounds = (k_sweeps + slots_max - 1) // slots_max
    return bands, (k_sweeps + rounds - 1) // rounds


def replay(rows, cols, nt, kb, k_sweeps, grid_max=148 * 8, publish_first=True, slots=None):
    """Returns the CTAs that never finish (empty list: no deadlock)."""
    cp = cols
    bands, gs = launch_shape(rows, cols, nt, kb, k_sweeps, grid_max)
    if slots is not None:
        gs = slots
    if gs < 1:
        return []
    ntasks = lambda s: (k_sweeps - s + gs - 1) // gs
    total = {(s, b): ntasks(s) * cp + nt - 1 for s in range(gs) for b in range(bands)}
    done = dict.fromkeys(total, 0)
    pub = dict.fromkeys(total, 0)
    moved = True
    while moved:
        moved = False
        for (s, b), tot in total.items():
            S = done[(s, b)]
            if S >= tot:
                continue
            if S % kb == 0:
                if publish_first:
                    pub[(s, b)] = S
                ps, adj = (gs - 1, cp) if s == 0 else (s - 1, 0)
                ok = True
                if gs > 1:
                    need = min(S + kb + 2 - adj, total[(ps, b)])
                    ok &= need <= 0 or pub[(ps, b)] >= need
                if b + 1 < bands:
                    need = min(S + kb + 2 - nt - adj, total[(ps, b + 1)])
                    ok &= need <= 0 or pub[(ps, b + 1)] >= need
                if b > 0:
                    ok &= pub[(s, b - 1)] >= min(S + kb + nt, tot)
                if not ok:
                    continue
                pub[(s, b)] = S
            done[(s, b)] = S + 1
            moved = True
            if S + 1 == tot:
                pub[(s, b)] = tot
    return [(k, done[k], total[k]) for k in total if done[k] < total[k]]


if __name__ == "__main__":
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    bad = 0
    for _ in range(300):
        nt, kb = rng.choice([64, 256, 512]), rng.choice([1, 2, 4, 8, 16])
        rows, cols, k = rng.randint(3, 1500), rng.randint(3 * kb + 4, 700), rng.randint(1, 120)
        stuck = replay(rows, cols, nt, kb, k)
        if stuck:
            bad += 1
            print("DEADLOCK", rows, cols, nt, kb, k, stuck[:3])
    print("the withheld-publication protocol at 67x300, 64-row bands, kb 8, 15 slots:",
          "deadlocks" if replay(67, 300, 64, 8, 90, publish_first=False, slots=15) else "runs")
    print("deadlocks:", bad)
    sys.exit(1 if bad else 0)
