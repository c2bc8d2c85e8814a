"""The driver mirrors (pl_inertial_slam_b200/drivers.py) against the same compositions evaluated with
the oracle: stereo matchGrid + gates, f2f match, KF<->KF / map->KF matchGrid with the stale-vector
match() fallback, loop-closure match + inlier gate."""
import numpy as np
import pytest

import oracle
from pl_inertial_slam_b200 import grid as G
from pl_inertial_slam_b200 import synth

pytestmark = pytest.mark.gpu
port = oracle.port
IW, IH = synth.INV_W, synth.INV_H


@pytest.fixture(scope="module")
def D(plm_lib):
    from pl_inertial_slam_b200 import drivers
    return drivers


def test_stereo_drivers(D):
    sp = synth.make_stereo_pair(synth.SEED0 + 1)
    res = D.matchStereoPoints(sp.kp_l, sp.kp_r, sp.pdesc_l, sp.pdesc_r, IW, IH)
    a = synth.stereo_points_grid_args(sp)
    n_o, m_o = port.match_grid_points(a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["rows"], a["cols"], a["d2"],
                                      a["win"], 0.9, True)
    _, k_o, d_o = port.stereo_filter_points(sp.kp_l, sp.kp_r, m_o)
    assert (res.matches_12 == m_o).all() and (res.keep == k_o).all() and (res.disparity == d_o).all()
    assert (res.desc_kept == sp.pdesc_l[k_o.astype(bool)]).all() and 200 < len(res.desc_kept) <= n_o

    resl = D.matchStereoLines(sp.ln_l, sp.ln_r, sp.ldesc_l, sp.ldesc_r, IW, IH)
    b = synth.stereo_lines_grid_args(sp)
    n_o, m_o = port.match_grid_lines(b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["rows"], b["cols"], b["d2"],
                                     b["dirs2"], 0.75, b["win"], 0.9, True)
    _, k_o, d_o = port.stereo_filter_lines(sp.ln_l, sp.ln_r, m_o)
    assert (resl.matches_12 == m_o).all() and (resl.keep == k_o).all()
    assert np.array_equal(resl.disparity, d_o, equal_nan=True)
    assert D.matchStereoPoints(sp.kp_l[:0], sp.kp_r, sp.pdesc_l[:0], sp.pdesc_r, IW, IH) is None


def test_f2f_and_loop_closure_drivers(D):
    prev, curr = synth.make_temporal_pair(synth.SEED0 + 2)
    m = D.matchF2FPoints(prev.pdesc_l, curr.pdesc_l)
    n_o, m_o = port.match(prev.pdesc_l, curr.pdesc_l, np.float32(0.9), True)
    assert (m == m_o).all()
    ml = D.matchF2FLines(prev.ldesc_l, curr.ldesc_l)
    n_l, m_lo = port.match(prev.ldesc_l, curr.ldesc_l, np.float32(0.9), True)
    assert (ml == m_lo).all()
    assert len(D.matchF2FPoints(prev.pdesc_l[:0], curr.pdesc_l)) == 0
    lc = D.loopClosureMatch(prev.pdesc_l, curr.pdesc_l, prev.ldesc_l, curr.ldesc_l)
    assert lc["common_pt"] == n_o and lc["common_ls"] == n_l
    assert lc["inl_ratio_condition"] == (100.0 * n_o / 600 > 30.0 and 100.0 * n_l / 200 > 30.0)


@pytest.mark.parametrize("force_fallback", [False, True])
def test_kf2kf_and_map2kf_drivers(D, force_fallback):
    prev, curr = synth.make_temporal_pair(synth.SEED0 + 2)
    if force_fallback:
        # projected coordinates far from the true positions: matchGrid finds < minPointMatches -> match()
        # runs on the vector matchGrid just filled
        pj = prev.kp_l.astype(np.float64) + 300.0
    else:
        pj = prev.kp_l.astype(np.float64)
    n_g, m_g = D.matchKF2KFPoints(pj, prev.pdesc_l, curr.kp_l, curr.pdesc_l, IW, IH)
    coords = np.stack([np.trunc(pj[:, 0] * IW), np.trunc(pj[:, 1] * IH)], 1).astype(np.int32)
    c = curr.kp_l.astype(np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * IW, c[:, 1] * IH)
    n_o, m_o = port.match_grid_points(coords, prev.pdesc_l, cs, ci, 48, 64, curr.pdesc_l, np.array([3, 3, 3, 3], np.int32), 0.9, True)
    if n_o < 10:
        n_o, m_o = port.match(prev.pdesc_l, curr.pdesc_l, np.float32(0.9), True, m12=m_o)
        assert force_fallback
    assert n_g == n_o and (m_g == m_o).all()

    # lines: the reference's pixel-coordinate quirk for the projected query lines
    pjl = prev.ln_l.astype(np.float64) + (300.0 if force_fallback else 0.0)
    n_g, m_g = D.matchKF2KFLines(pjl, prev.ldesc_l, curr.ln_l, curr.ldesc_l, IW, IH)
    cs, ci, dirs = D.line_grid(curr.ln_l, IW, IH)
    coords = np.trunc(pjl).astype(np.int32)
    n_o, m_o = port.match_grid_lines(coords, prev.ldesc_l, cs, ci, 48, 64, curr.ldesc_l, dirs, 0.75,
                                     np.array([3, 3, 3, 3], np.int32), 0.9, True)
    if n_o < 6:
        n_o, m_o = port.match(prev.ldesc_l, curr.ldesc_l, np.float32(0.9), True, m12=m_o)
    assert n_g == n_o and (m_g == m_o).all()

    # map -> keyframe (config 4 shape, small): 5000 map points against the frame's points
    mdesc, mxy = synth.make_map_points(synth.SEED0 + 4, 5000, curr)
    pjm = (mxy.astype(np.float64) + 0.5) / np.array([IW, IH]) + (5000.0 if force_fallback else 0.0)
    n_g, m_g = D.matchMap2KFPoints(pjm, mdesc, curr.kp_l, curr.pdesc_l, IW, IH)
    coords = np.stack([np.trunc(pjm[:, 0] * IW), np.trunc(pjm[:, 1] * IH)], 1).astype(np.int32)
    c = curr.kp_l.astype(np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * IW, c[:, 1] * IH)
    n_o, m_o = port.match_grid_points(coords, mdesc, cs, ci, 48, 64, curr.pdesc_l, np.array([3, 3, 3, 3], np.int32), 0.9, True)
    if n_o < 10:
        n_o, m_o = port.match(mdesc, curr.pdesc_l, np.float32(0.9), True, m12=m_o)
    assert n_g == n_o and (m_g == m_o).all()
