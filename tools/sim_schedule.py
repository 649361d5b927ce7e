#!/usr/bin/env python
"""Padding of conflict-free sweep schedules, simulated on the CPU (numpy).

A (warp, panel) block of the sweep layout takes as many steps as its worst
quarter warp needs; a quarter warp with O owners and C shared-memory bank classes needs
Delta = max(row sums, column sums) of its O x C count matrix (Koenig).  This script draws
per-owner entries like bench.py's synthetic matrix (n_per_owner uniform rows out of n_other,
cut into panels of `panel_rows`), and reports pad entries / real entries for

  pairs : 16 owners per warp (one per lane pair), quarter warp = 4 owners x 4 classes   [current]
  lanes : 32 owners per warp (one per lane),      quarter warp = 8 owners x 8 classes   [candidate
          for small K: halves the per-entry instruction overhead, needs rows with
          (ST/2) odd so that row mod 8 is the bank class of a 16-byte unit]

both with the optimal step count and, for `lanes`, with the greedy matching a device kernel would
run (8! permutations cannot be enumerated per step like the 24 of the 4 x 4 case).
"""
import argparse

import numpy as np


def greedy_steps(n):
    """Steps a greedy colouring needs for count matrix n (owners x classes): every step matches
    owners to distinct classes, critical rows / columns (remaining degree == steps left) first."""
    n = n.copy()
    O, C = n.shape
    steps = 0
    while n.sum() > 0:
        rs, cs = n.sum(1), n.sum(0)
        used_c = np.zeros(C, bool)
        # owners by remaining degree (largest first); each takes its fullest free class,
        # preferring classes that are themselves critical
        for o in np.argsort(-rs, kind="stable"):
            if rs[o] == 0:
                continue
            cand = np.where((n[o] > 0) & ~used_c)[0]
            if cand.size == 0:
                continue
            c = cand[np.argmax(cs[cand] * 1000 + n[o, cand])]
            used_c[c] = True
            n[o, c] -= 1
        steps += 1
    return steps


def greedy_place(lists, classes):
    """Reference of the placement a one-lane-per-owner layout kernel has to produce for ONE
    (quarter warp, panel): `lists[o][c]` = panel-local rows of owner o in bank class c (each list
    is consumed in order).  Returns schedule[step][o] = row or -1 (pad): per step every class is
    read by at most one owner, every entry appears exactly once, and the number of steps is what
    `greedy_steps` reports for the count matrix (both apply the same rule)."""
    O = len(lists)
    n = np.array([[len(lists[o][c]) for c in range(classes)] for o in range(O)], dtype=np.int64)
    used = np.zeros_like(n)
    schedule = []
    while n.sum() > 0:
        rs, cs = n.sum(1), n.sum(0)
        taken = np.zeros(classes, bool)
        step = [-1] * O
        for o in np.argsort(-rs, kind="stable"):
            if rs[o] == 0:
                continue
            cand = np.where((n[o] > 0) & ~taken)[0]
            if cand.size == 0:
                continue
            c = cand[np.argmax(cs[cand] * 1000 + n[o, cand])]
            taken[c] = True
            step[o] = lists[o][c][used[o, c]]
            used[o, c] += 1
            n[o, c] -= 1
        schedule.append(step)
    return schedule


def simulate(owners_per_warp, per_quarter, classes, n_other, panel_rows, n_per_owner, n_warps, rng, greedy):
    npanel = (n_other + panel_rows - 1) // panel_rows
    real = pad_opt = pad_greedy = 0
    for _ in range(n_warps):
        # rows of every owner: n_per_owner distinct uniform draws
        cnt = np.zeros((owners_per_warp, npanel, classes), dtype=np.int64)
        for o in range(owners_per_warp):
            rows = rng.choice(n_other, size=n_per_owner, replace=False)
            np.add.at(cnt[o], (rows // panel_rows, (rows % panel_rows) % classes), 1)
        for p in range(npanel):
            d_opt = d_gr = 0
            for q0 in range(0, owners_per_warp, per_quarter):
                m = cnt[q0:q0 + per_quarter, p, :]
                d_opt = max(d_opt, int(max(m.sum(1).max(), m.sum(0).max())))
                if greedy:
                    d_gr = max(d_gr, greedy_steps(m))
            d_opt += d_opt & 1                   # two steps per stream element
            d_gr += d_gr & 1
            tot = int(cnt[:, p, :].sum())
            real += tot
            pad_opt += owners_per_warp * d_opt - tot
            pad_greedy += owners_per_warp * d_gr - tot
    return pad_opt / real, (pad_greedy / real if greedy else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warps", type=int, default=24)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    print("cells own, genes stream: 1903 nonzeros per cell out of 20000 genes (cfg-3)")
    for rows, what in ((1448, "K=20"), (2416, "K=7/10")):
        a, _ = simulate(16, 4, 4, 20000, rows, 1903, args.warps, rng, greedy=False)
        b, g = simulate(32, 8, 8, 20000, rows, 1903, max(args.warps // 4, 2), rng, greedy=True)
        print("  panel %4d rows (%s): pairs 4x4 optimal %.1f %%   lanes 8x8 optimal %.1f %%, greedy %.1f %%"
              % (rows, what, 100 * a, 100 * b, 100 * g))


if __name__ == "__main__":
    main()
