"""tools/sim_schedule.py is the CPU model of the sweep layout's conflict-free schedule
(DESIGN.md §3, §7): it must reproduce the padding the device layout was measured to have, and
its greedy 8-class placement -- the reference for the one-lane-per-owner layout planned for
small K -- must be a valid schedule (integer work: exact checks)."""
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import sim_schedule as sim          # noqa: E402


def test_simulated_padding_matches_the_measured_layout():
    # bench.py's cfg-3 layout: padded / real = 1.168 (cells own, 1448-row panels; profiles/r1b_bench_n1.json)
    rng = np.random.default_rng(0)
    pad, _ = sim.simulate(16, 4, 4, 20000, 1448, 1903, 8, rng, greedy=False)
    assert abs(pad - 0.168) < 0.012
    rng = np.random.default_rng(0)
    pad, _ = sim.simulate(16, 4, 4, 20000, 2416, 1903, 8, rng, greedy=False)      # K = 7 / 10: 1.130
    assert abs(pad - 0.130) < 0.012


def test_greedy_eight_class_placement_is_a_valid_schedule():
    rng = np.random.default_rng(3)
    worst = 0.0
    for trial in range(40):
        O = C = 8
        rows_per_owner = rng.integers(0, 60, size=O)
        if trial == 0:
            rows_per_owner[:] = 0                        # an empty quarter warp
        if trial == 1:
            rows_per_owner[:] = [200, 0, 0, 0, 0, 0, 0, 1]   # one heavy owner
        lists = []
        for o in range(O):
            rows = np.sort(rng.choice(2416, size=int(rows_per_owner[o]), replace=False))
            lists.append([[int(r) for r in rows if r % C == c] for c in range(C)])
        counts = np.array([[len(lists[o][c]) for c in range(C)] for o in range(O)])
        sched = sim.greedy_place(lists, C)
        # every entry exactly once, each owner's lists consumed in order
        for o in range(O):
            got = [s[o] for s in sched if s[o] >= 0]
            assert sorted(got) == sorted(r for c in range(C) for r in lists[o][c])
            for c in range(C):
                assert [r for r in got if r % C == c] == lists[o][c]
        # no two owners of the quarter warp read the same bank class in a step
        for s in sched:
            cls = [r % C for r in s if r >= 0]
            assert len(cls) == len(set(cls))
        # same step count as the counting version, never below the Koenig bound
        delta = int(max(counts.sum(1).max(), counts.sum(0).max())) if counts.sum() else 0
        assert len(sched) == sim.greedy_steps(counts) >= delta
        if delta:
            worst = max(worst, len(sched) / delta)
    assert worst <= 1.10                                 # greedy stays within 10 % of optimal even on ragged inputs
