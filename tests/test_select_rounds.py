"""The scheme sr_select_kernel uses instead of the reference's serial greedy walk (SR:559-675), restated in numpy / plain
Python and checked against the oracle's walk on whole sweeps (CPU only; the kernel itself is checked bit for bit by the
GPU tests):

  sort  -> ranks = position in the stable sort of (curvature, index);
  walk  -> a candidate (curvature > 0.1 for the sharp walk, < 0.1 for the flat one, not marked) is picked iff none of its
           DOMINATORS is picked: the neighbours within +-5 that come earlier in the walk and whose suppression span
           (SR:597-622, cut at a gap > 0.05) covers it.  Resolved in rounds (UNDECIDED -> IN when no dominator is IN or
           UNDECIDED, -> OUT when one is IN); the count limits (16 + 4 sharp, 32 flat) keep a prefix in walk order and
           only the kept picks (the 32nd flat one excepted, SR:635-638) leave marks for the next walk.
"""
import numpy as np
import pytest


def _gap_flags(full):
    """gap[i]: squared distance p_i - p_(i-1) > 0.05, in the reference's float32 operation order (SR:604, 617)."""
    p = full[:, :3].astype(np.float32)
    d = p[1:] - p[:-1]
    g2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    gap = np.zeros(full.shape[0], bool)
    gap[1:] = g2.astype(np.float64) > 0.05
    return gap


def _reach(gap, i, n):
    nf = 0
    for l in range(1, 6):
        if i + l >= n or gap[i + l]:
            break
        nf = l
    nb = 0
    for l in range(1, 6):
        if i - l < 0 or gap[i - l + 1]:
            break
        nb = l
    return nf, nb


def select_by_rounds(curv, picked0, gap, scan_start, scan_end, n):
    """Labels of all rings by the rounds scheme; returns (label, total rounds, pick order per ring)."""
    label = np.zeros(n, np.int32)
    picked = picked0.astype(bool).copy()
    rounds = 0
    order = []
    for S, E in zip(scan_start, scan_end):
        ring_order = {"sharp": [], "less": [], "flat": []}
        for j in range(6):
            sp = (S * (6 - j) + E * j) // 6
            ep = (S * (5 - j) + E * (j + 1)) // 6 - 1
            m = ep - sp + 1
            if m <= 0:
                continue
            cv = curv[sp:ep + 1]
            rank = np.empty(m, np.int64)
            rank[np.argsort(cv, kind="stable")] = np.arange(m)
            reach = [_reach(gap, sp + t, n) for t in range(m)]
            for kind in (0, 1):
                passes = (cv.astype(np.float64) > 0.1) if kind == 0 else (cv.astype(np.float64) < 0.1)
                prio = -rank if kind == 0 else rank  # smaller = earlier in the walk
                UND, IN, OUT = 1, 2, 0
                state = np.where(passes & ~picked[sp:ep + 1], UND, OUT)
                dom = []
                for t in range(m):
                    ds = []
                    for d in range(1, 6):
                        lo, hi = t - d, t + d
                        if lo >= 0 and reach[lo][0] >= d and prio[lo] < prio[t]:
                            ds.append(lo)
                        if hi < m and reach[hi][1] >= d and prio[hi] < prio[t]:
                            ds.append(hi)
                    dom.append(ds)
                while (state == UND).any():
                    rounds += 1
                    prev = state.copy()
                    for t in np.nonzero(prev == UND)[0]:
                        st = [prev[x] for x in dom[t]]
                        if IN in st:
                            state[t] = OUT
                        elif UND not in st:
                            state[t] = IN
                    assert (state != prev).any()  # the best-ranked undecided candidate always resolves
                ins = sorted(np.nonzero(state == IN)[0], key=lambda t: prio[t])
                keep, marking = (20, 20) if kind == 0 else (32, 31)
                for num, t in enumerate(ins[:keep], 1):
                    i = sp + t
                    if kind == 0:
                        label[i] = 2 if num <= 16 else 1
                        if num <= 16:
                            ring_order["sharp"].append(i)
                        ring_order["less"].append(i)
                    else:
                        label[i] = -1
                        ring_order["flat"].append(i)
                    if num <= marking:
                        nf, nb = reach[t]
                        picked[i - nb:i + nf + 1] = True
        order.append(ring_order)
    return label, rounds, order


def _dense_cylinder():
    rng = np.random.default_rng(11)
    per_ring, angles = 1300, np.array([-15, -13, -11, -9, -7, -5, -4, -3, -2, -1, 0, 1, 3, 5, 7, 9], np.float64)
    az = np.linspace(0.0, 2 * np.pi, per_ring, endpoint=False)
    rows = []
    for k in range(per_ring):
        for e in angles:
            rad = 8.0 + 2.5 * np.sin(3 * az[k]) + (0.6 if (k // 37) % 5 == 0 else 0.0)
            rad = 12.0 if e > 4 else rad + rng.normal(0, 0.01)  # the top rings: a noise-free cylinder -> curvature ties
            rows.append((rad * np.cos(-az[k]), rad * np.sin(-az[k]), rad * np.tan(np.deg2rad(e))))
    return np.asarray(rows, np.float32)


@pytest.mark.parametrize("case", ["scene", "cylinder_with_ties"])
def test_rounds_scheme_equals_the_serial_walk(orc, sweeps16, case):
    xyz = sweeps16[3] if case == "scene" else _dense_cylinder()
    sr = orc.ScanRegistration()
    ref = sr.extract(xyz)
    full = ref["full"]
    n = full.shape[0]
    curv = sr.curvature()[:n]
    start, end = sr.ints("scan_start"), sr.ints("scan_end")
    assert (start[1:] > 0).all()  # no virtual rings in these sweeps (those are replayed serially on the device as well)
    label, rounds, order = select_by_rounds(curv, sr.ints("picked_mask")[:n], _gap_flags(full), start, end, n)
    want = sr.ints("label")[:n]
    assert np.array_equal(label[5:n - 5], want[5:n - 5])
    # pick order = cloud order (ring, sector, walk): the feature clouds are the full cloud's rows in that order
    for name, key in (("sharp", "sharp"), ("less_sharp", "less"), ("flat", "flat")):
        idx = np.array([i for ring in order for i in ring[key]], np.int64)
        assert np.array_equal(full[idx].view(np.uint32), ref[name].view(np.uint32)), name
    assert rounds < 16 * 12 * 12  # a handful of rounds per walk, not one per pick
    if case == "cylinder_with_ties":
        assert len(np.unique(curv[start[15]:end[15]])) < (end[15] - start[15]) // 2  # the ties are really there
