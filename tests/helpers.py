"""Shared case builders for the parity tests (same seeded inputs go to the oracle and the GPU)."""
import numpy as np

from pl_inertial_slam_b200 import grid as G
from pl_inertial_slam_b200 import synth


def random_grid_case(rng, n1, n2, rows=48, cols=64, is_lines=False, tie=False, win=(3, 3, 3, 3), off_grid=0.1,
                     bad_items=0, zero_len=0):
    """A matchGrid case with uniformly scattered features; returns kwargs for oracle / GPU."""
    d1 = synth.tie_stress_desc(rng, n1) if tie else synth.rand_desc(rng, n1)
    d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
    if not tie and n1 and n2:
        k = min(n1, n2) // 2
        src = rng.integers(0, n2, k)
        dst = rng.choice(n1, k, replace=False)
        d1[dst] = synth.flip_bits(rng, d2[src], 0.08)
    lo_x, hi_x = (-3, cols + 3) if off_grid else (0, cols)
    lo_y, hi_y = (-3, rows + 3) if off_grid else (0, rows)
    if not is_lines:
        xy = np.stack([rng.integers(lo_x, hi_x, n1), rng.integers(lo_y, hi_y, n1)], 1).astype(np.int32)
        px = rng.uniform(-1.5 if off_grid else 0, cols + (1.5 if off_grid else 0), n2)
        py = rng.uniform(-1.5 if off_grid else 0, rows + (1.5 if off_grid else 0), n2)
        cs, ci = G.csr_from_points(px, py, rows, cols)
        if bad_items and len(ci):
            ci = ci.copy()
            pos = rng.integers(0, len(ci), bad_items)
            ci[pos] = rng.choice([-1, -7, n2, n2 + 5], bad_items)
        return dict(coords=xy, d1=d1, cell_start=cs, cell_items=ci, rows=rows, cols=cols, d2=d2,
                    win=np.array(win, np.int32), dirs2=None)
    s = np.stack([rng.integers(lo_x, hi_x, n1), rng.integers(lo_y, hi_y, n1)], 1)
    e = s + rng.integers(-8, 9, (n1, 2))
    if zero_len and n1:
        z = rng.choice(n1, min(zero_len, n1), replace=False)
        e[z] = s[z]  # zero-length query -> NaN direction -> passes the direction filter
    xyxy = np.concatenate([s, e], 1).astype(np.int32)
    a = np.stack([rng.uniform(0, cols, n2), rng.uniform(0, rows, n2)], 1)
    b = a + rng.uniform(-9, 9, (n2, 2))
    cs, ci = G.csr_from_lines(a[:, 0], a[:, 1], b[:, 0], b[:, 1], rows, cols)
    dirs2 = G.line_directions(a[:, 0], a[:, 1], b[:, 0], b[:, 1])
    if zero_len and n2:
        z = rng.choice(n2, min(zero_len, n2), replace=False)
        dirs2[z] = np.nan
    return dict(coords=xyxy, d1=d1, cell_start=cs, cell_items=ci, rows=rows, cols=cols, d2=d2,
                win=np.array(win, np.int32), dirs2=dirs2)


def oracle_grid(port, case, ratio, best_lr, line_sim_th=0.75, m12=None):
    if case["dirs2"] is None:
        return port.match_grid_points(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                      case["rows"], case["cols"], case["d2"], case["win"], ratio, best_lr, m12)
    return port.match_grid_lines(case["coords"], case["d1"], case["cell_start"], case["cell_items"], case["rows"],
                                 case["cols"], case["d2"], case["dirs2"], line_sim_th, case["win"], ratio, best_lr,
                                 m12)


def gpu_grid(case, ratio, best_lr, line_sim_th=0.75, m12=None, ctx=None):
    from pl_inertial_slam_b200 import matching as M
    M.Config.bestLRMatches = bool(best_lr)
    M.Config.minRatio12P = ratio
    M.Config.lineSimTh = line_sim_th
    n1 = case["d1"].shape[0]
    buf = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
    grid = (case["cell_start"], case["cell_items"], case["rows"], case["cols"])
    try:
        if case["dirs2"] is None:
            n = M.matchGrid(case["coords"], case["d1"], grid, case["d2"], case["win"], buf, ctx=ctx)
        else:
            n = M.matchGrid(case["coords"], case["d1"], grid, case["d2"], case["dirs2"], case["win"], buf, ctx=ctx)
    finally:
        M.Config.bestLRMatches = True
        M.Config.minRatio12P = 0.9
        M.Config.lineSimTh = 0.75
    return n, buf
