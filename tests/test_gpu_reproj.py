"""CUDA local-map selection / reprojection gates (csrc/plm_reproj.cuh, plm_map_select / plm_map_gate) against the
restatement of mapHandler.cpp:583-682, :685-803 -- bit for bit, fp64 outputs by their bit patterns -- and the end to end
matchMap2KF* drivers (selection -> matchGrid -> match fallback -> gate) against the oracle pieces chained the same way."""
import numpy as np
import pytest

import oracle
from helpers import oracle_grid
from pl_inertial_slam_b200 import grid as G
from pl_inertial_slam_b200 import synth
from test_reproj import CAM, H, INV_H, INV_W, W, make_scene

pytestmark = pytest.mark.gpu
port = oracle.port


@pytest.fixture(scope="module")
def D(plm_lib):
    from pl_inertial_slam_b200 import drivers
    return drivers


@pytest.mark.parametrize("lines", [False, True])
@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 70_001])
def test_select_and_gate_vs_restatement(D, lines, n):
    T, X, active = make_scene(100 + n % 97, n, lines)
    view = D.map_view(T, CAM, INV_W, INV_H, W, H)
    for act in (active, None):
        sel, coords, pf = D.mapSelect(X, act, view)
        s_o, c_o, p_o = port.map_select(X, act, T, CAM, INV_W, INV_H, W, H)
        assert np.array_equal(sel, s_o) and np.array_equal(coords, c_o)
        assert np.array_equal(pf.view(np.uint64), p_o.view(np.uint64))
    if len(sel) == 0:
        return
    rng = np.random.default_rng(n)
    n2 = 300
    m12 = rng.integers(-1, n2, len(sel)).astype(np.int32)
    if not lines:
        feat = rng.uniform(0, 700, (n2, 2))
        idx = np.flatnonzero(m12 >= 0)[::2]
        feat_idx = m12[idx]
        pf2 = pf.copy()
        pf2[idx] = feat[feat_idx] + rng.normal(0, 0.7, (len(idx), 2))
    else:
        feat = rng.normal(0, 1, (n2, 3))
        feat[:, 2] = -(feat[:, 0] * 350 + feat[:, 1] * 240) + rng.normal(0, 2, n2)
        pf2 = pf
    c0 = int((m12 >= 0).sum())
    c_g, ok_g = D.mapGate(pf2, m12, feat, 1.0, c0)
    c_o, ok_o = port.map_gate(pf2, m12, feat, 1.0, c0)
    assert c_g == c_o and np.array_equal(ok_g, ok_o)


def test_match_map2kf_points_end_to_end(D):
    """A local map of 40 000 landmarks, a keyframe with 500 unmatched stereo points that observe some of them."""
    rng = np.random.default_rng(7)
    n, n2 = 40_000, 500
    T, X, active = make_scene(5, n)
    view = D.map_view(T, CAM, INV_W, INV_H, W, H)
    med = synth.rand_desc(rng, n)
    s_o, c_o, p_o = port.map_select(X, active, T, CAM, INV_W, INV_H, W, H)
    seen = rng.choice(len(s_o), n2, replace=False)
    pl = p_o[seen] + rng.normal(0, 0.8, (n2, 2))                       # observed where the landmark projects, +- noise
    pl = np.clip(pl, 1.0, [W - 2.0, H - 2.0])
    d2 = synth.flip_bits(rng, med[s_o[seen]], 0.06)
    d2[::9] = synth.rand_desc(rng, len(d2[::9]))                       # some keyframe points see nothing in the map
    out = D.matchMap2KFPointsFull(X, active, med, view, pl, d2)
    # the same chain through the oracle
    cs, ci = G.csr_from_points(pl[:, 0] * INV_W, pl[:, 1] * INV_H)
    case = dict(coords=c_o, d1=np.ascontiguousarray(med[s_o]), cell_start=cs, cell_items=ci, rows=G.GRID_ROWS, cols=G.GRID_COLS,
                d2=d2, win=np.array([3, 3, 3, 3], np.int32), dirs2=None)
    n_o, m_o = oracle_grid(port, case, 0.9, 1)
    assert n_o >= 10                                                   # no fallback on this case
    cnt_o, ok_o = port.map_gate(p_o, m_o, pl, 1.0, n_o)
    assert np.array_equal(out["sel"], s_o) and np.array_equal(out["m12"], m_o)
    assert np.array_equal(out["ok"], ok_o) and out["matches"] == cnt_o and 50 < cnt_o <= n_o
