"""CPU tests: the oracle restatement against (1) golden vectors generated from the reference itself
(tools/make_golden.py), (2) the compiled reference when oracle/_ref/libplref.so exists, (3) OpenCV's
own BFMatcher (cv2 wheel) for the one un-vendored dependency, and the host-side grid logic."""
import glob
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT
from helpers import oracle_grid, random_grid_case
from pl_inertial_slam_b200 import grid as G
from pl_inertial_slam_b200 import synth

port, ref = oracle.port, oracle.ref
GOLDEN = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libplref.so not built")


def _golden(prefix):
    files = sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))
    assert files, "golden fixtures missing"
    return files


@pytest.mark.parametrize("path", _golden("brute"))
def test_port_vs_golden_brute(path):
    z = np.load(path)
    for nnr in (0.75, 0.9):
        for blr in (0, 1):
            n, m = port.match(z["d1"], z["d2"], nnr, blr)
            assert n == int(z[f"n_{nnr}_{blr}"]) and (m == z[f"m_{nnr}_{blr}"]).all()
            n, m = port.match(z["d1"], z["d2"], nnr, blr, m12=z["stale"])
            assert n == int(z[f"ns_{nnr}_{blr}"]) and (m == z[f"ms_{nnr}_{blr}"]).all()


def _case_from_npz(z):
    return dict(coords=z["coords"], d1=z["d1"], cell_start=z["cell_start"], cell_items=z["cell_items"],
                rows=int(z["rows"]), cols=int(z["cols"]), d2=z["d2"], win=z["win"],
                dirs2=z["dirs2"] if int(z["is_lines"]) else None)


@pytest.mark.parametrize("path", _golden("grid"))
def test_port_vs_golden_grid(path):
    z = np.load(path)
    case = _case_from_npz(z)
    for ratio in (0.75, 0.9, 1.0):
        for blr in (0, 1):
            n, m = oracle_grid(port, case, ratio, blr)
            assert n == int(z[f"n_{ratio}_{blr}"]) and (m == z[f"m_{ratio}_{blr}"]).all()


def test_line_coords_vs_golden():
    z = np.load(os.path.join(GOLDEN, "line_coords.npz"))
    seg, cells, offs = z["seg"], z["cells"], z["offs"]
    for i, s in enumerate(seg):
        want = cells[offs[i]:offs[i + 1]]
        assert (port.line_coords(*s) == want).all()
        assert (np.array(G.getLineCoords(*s), np.int32).reshape(-1, 2) == want).all()
    ids, cx, cy = G.line_cells(seg[:, 0], seg[:, 1], seg[:, 2], seg[:, 3])
    assert (np.stack([cx, cy], 1) == cells).all()
    assert (np.bincount(ids, minlength=len(seg)) == np.diff(offs)).all()


def test_distance_known_answers():
    z = np.zeros(32, np.uint8)
    f = np.full(32, 255, np.uint8)
    assert port.distance(z, z) == 0 and port.distance(z, f) == 256
    one = z.copy(); one[17] = 0x10
    assert port.distance(z, one) == 1
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = synth.rand_desc(rng, 2)
        assert port.distance(a, b) == int(np.unpackbits(a ^ b).sum())


def test_knn2_vs_cv2():
    """cv::BFMatcher is the one dependency not under /root/reference (OpenCV 3.3 there, 4.13 here)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for n1, n2, tie in [(50, 2, 1), (70, 3, 1), (100, 33, 1), (64, 1025, 1), (300, 300, 0)]:
        d1 = synth.tie_stress_desc(rng, n1) if tie else synth.rand_desc(rng, n1)
        d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
        idx, dist = port.knn2(d1, d2)
        m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
        assert (np.array([[x.trainIdx for x in r] for r in m]) == idx).all()
        assert (np.array([[int(x.distance) for x in r] for r in m]) == dist).all()


@needs_ref
def test_port_vs_reference_random():
    rng = np.random.default_rng(2)
    for _ in range(30):
        n1, n2 = int(rng.integers(2, 200)), int(rng.integers(2, 200))
        tie = bool(rng.integers(0, 2))
        d1 = synth.tie_stress_desc(rng, n1) if tie else synth.rand_desc(rng, n1)
        d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
        for blr in (0, 1):
            for lr_par in (0, 1):
                assert_same(port.match(d1, d2, 0.9, blr), ref.match(d1, d2, 0.9, best_lr=blr, lr_parallel=lr_par))
        a, b = d1[0], d2[0]
        assert port.distance(a, b) == ref.distance(a, b)


def assert_same(x, y):
    assert x[0] == y[0] and (x[1] == y[1]).all()


@needs_ref
@pytest.mark.parametrize("is_lines", [False, True])
def test_port_vs_reference_grid_random(is_lines):
    """The reference iterates an unordered_set; the port iterates cells in order -- 200 random
    cases with heavy ties show the outputs agree for ratio <= 1 (SURVEY 8a note 1)."""
    rng = np.random.default_rng(3 + int(is_lines))
    for it in range(100):
        n1, n2 = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        w = tuple(int(v) for v in rng.integers(0, 6, 4))
        case = random_grid_case(rng, n1, n2, rows=int(rng.integers(1, 20)), cols=int(rng.integers(1, 20)),
                                is_lines=is_lines, tie=bool(it % 2), win=w, bad_items=int(rng.integers(0, 3)),
                                zero_len=int(rng.integers(0, 3)))
        for ratio in (0.75, 1.0):
            for blr in (0, 1):
                got = oracle_grid(port, case, ratio, blr)
                if is_lines:
                    want = ref.match_grid_lines(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                                case["rows"], case["cols"], case["d2"], case["dirs2"], 0.75,
                                                case["win"], ratio, blr)
                else:
                    want = ref.match_grid_points(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                                 case["rows"], case["cols"], case["d2"], case["win"], ratio, blr)
                assert_same(got, want)


def test_grid_mirror_matches_csr_builders():
    """GridStructure.at/get/to_csr (host mirror) against the vectorised CSR builders."""
    rng = np.random.default_rng(5)
    px, py = rng.uniform(-2, 66, 400), rng.uniform(-2, 50, 400)
    g = G.GridStructure(G.GRID_ROWS, G.GRID_COLS)
    for i in range(400):
        g.at(px[i], py[i]).append(i)
    cs, ci = g.to_csr()
    cs2, ci2 = G.csr_from_points(px, py)
    assert (cs == cs2).all() and (ci == ci2).all()
    w = G.GridWindow((3, 1), (0, 2))
    for (x, y) in [(0, 0), (63, 47), (10, 20), (-1, 5), (70, 50)]:
        s = set()
        g.get(x, y, w, s)
        want = set()
        for x_ in range(max(0, x - 3), min(64, x + 2)):
            for y_ in range(max(0, y), min(48, y + 3)):
                c = x_ * 48 + y_
                want.update(ci[cs[c]:cs[c + 1]].tolist())
        assert s == want
    seg = rng.uniform(0, 60, (50, 4))
    g2 = G.GridStructure(48, 64)
    for i, s in enumerate(seg):
        for (x, y) in G.getLineCoords(*s):
            g2.at(x, y).append(i)
    a, b = g2.to_csr()
    a2, b2 = G.csr_from_lines(seg[:, 0], seg[:, 1], seg[:, 2], seg[:, 3])
    assert (a == a2).all() and (b == b2).all()
    with pytest.raises(RuntimeError):
        G.GridStructure(48, 0)


def test_stereo_filter_port_properties():
    sp = synth.make_stereo_pair(synth.SEED0 + 1)
    m = np.arange(600, dtype=np.int32)
    n, keep, disp = port.stereo_filter_points(sp.kp_l, sp.kp_r, m)
    dy = np.abs(sp.kp_l[:, 1] - sp.kp_r[:, 1])
    dx = (sp.kp_l[:, 0] - sp.kp_r[:, 0]).astype(np.float64)
    want = (dy.astype(np.float64) <= 1.0) & (dx >= 1.0)
    assert (keep.astype(bool) == want).all() and n == want.sum()
    assert port.lib.plo_line_overlap_stereo(0.0, 10.0, 0.0, 10.0, 0.1) == 1.0
    assert port.lib.plo_line_overlap_stereo(0.0, 10.0, 20.0, 30.0, 0.1) == 0.0
    assert port.lib.plo_line_overlap_stereo(5.0, 5.05, 20.0, 30.0, 0.1) == 1.0  # horizontal: untouched


def _np_min(a, b):   # std::min(a, b) = (b < a) ? b : a
    return np.where(b < a, b, a)


def _np_max(a, b):   # std::max(a, b) = (a < b) ? b : a
    return np.where(a < b, b, a)


def numpy_stereo_line_gates(ln_l, ln_r, m12, min_disp, horiz_th, overlap_th, ratio_th):
    """Independent vectorised form of stereoFrame.cpp:359-385 + :416-426 + :484-519 (float64, numpy never contracts to
    FMA): endpoint interpolation with the reference's quirk (sp_r is overwritten before ep_r is interpolated)."""
    n2 = len(ln_r)
    ok = (m12 >= 0) & (m12 < n2)
    j = np.where(ok, m12, 0)
    spl_x, spl_y, epl_x, epl_y = ln_l.astype(np.float64).T
    spr_x, spr_y, epr_x, epr_y = ln_r.astype(np.float64)[j].T
    with np.errstate(all="ignore"):
        # lineSegmentOverlapStereo on the y coordinates
        sln, eln = _np_min(spl_y, epl_y), _np_max(spl_y, epl_y)
        spn, epn = _np_min(spr_y, epr_y), _np_max(spr_y, epr_y)
        length = eln - spn
        inner = np.where((epn > eln) & (spn < sln), eln - sln, _np_min(eln, epn) - _np_max(sln, spn))
        ov = np.where((epn < sln) | (spn > eln), 0.0, inner)
        ov = np.where(length > np.float64(np.float32(0.01)), ov / length, 0.0)
        ov = np.where(ov > 1.0, 1.0, ov)
        overlap = np.where(np.abs(epl_y - spl_y) > horiz_th, ov, 1.0)
        # endpoint interpolation: sp_r first, then ep_r with the UPDATED sp_r
        sx = (spr_x * (spl_y - epr_y) + epr_x * (spr_y - spl_y)) / (spr_y - epr_y)
        sy = spl_y
        ex = (sx * (epl_y - epr_y) + epr_x * (sy - epl_y)) / (sy - epr_y)
        ey = epl_y
        disp_s, disp_e = spl_x - sx, epl_x - ex
        bad = _np_min(disp_s, disp_e) / _np_max(disp_s, disp_e) < ratio_th
        disp_s, disp_e = np.where(bad, -1.0, disp_s), np.where(bad, -1.0, disp_e)
        keep = ok & (disp_s >= min_disp) & (disp_e >= min_disp) & (np.abs(spl_y - epl_y) > horiz_th) \
            & (np.abs(sy - ey) > horiz_th) & (overlap > overlap_th)
    return keep.astype(np.uint8), np.where(ok, disp_s, 0.0), np.where(ok, disp_e, 0.0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_stereo_line_gates_port_vs_numpy(seed):
    """plo_stereo_filter_lines (restated from stereoFrame.cpp, which cannot be compiled here) against an independent
    numpy form, bit for bit on the disparities."""
    sp = synth.make_stereo_pair(synth.SEED0 + 30 + seed)
    rng = np.random.default_rng(seed)
    n1, n2 = len(sp.ln_l), len(sp.ln_r)
    m12 = np.arange(n1, dtype=np.int32) % n2
    m12[rng.random(n1) < 0.2] = -1
    wrong = rng.random(n1) < 0.2
    m12[wrong] = rng.integers(0, n2, int(wrong.sum()))            # wrong partners: every reject branch occurs
    ln_l, ln_r = sp.ln_l.copy(), sp.ln_r.copy()
    ln_l[:10, 3] = ln_l[:10, 1] + 0.05                            # near-horizontal segments
    n, keep, disp = port.stereo_filter_lines(ln_l, ln_r, m12, 1.0, 0.1, 0.75, 0.7)
    k2, ds, de = numpy_stereo_line_gates(ln_l, ln_r, m12, 1.0, 0.1, 0.75, 0.7)
    assert np.array_equal(keep, k2) and n == int(k2.sum()) and 0 < n < (m12 >= 0).sum()
    fin = np.isfinite(ds) & np.isfinite(de)
    assert np.array_equal(disp[fin, 0].view(np.uint64), ds[fin].view(np.uint64))
    assert np.array_equal(disp[fin, 1].view(np.uint64), de[fin].view(np.uint64))
    assert np.array_equal(np.isnan(disp[:, 0]), np.isnan(ds))


def test_port_stereo_drivers_vs_reference_golden():
    """tests/golden/stereo.npz = outputs of the reference's own stereoFrame.cpp (tools/make_golden.py stereo): the
    restatement must reproduce them bit for bit also where the reference build is not available."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stereo.npz"))
    fx, _, cx, cy, b = z["cam_ref"]
    cam = np.array([b, fx, cx, cy])
    w, h = (int(x) for x in z["img_wh"])
    bits = lambda a: np.ascontiguousarray(a, np.float64).view(np.uint64)  # noqa: E731
    for c in range(int(z["n_cfg"])):
        g = lambda k, d: float(z[f"cfg{c}_{k}"]) if f"cfg{c}_{k}" in z.files else d  # noqa: E731
        for f in range(int(z["n_frames"])):
            d = {k: z[f"f{f}_{k}"] for k in ("kp_l", "kp_r", "pdesc_l", "pdesc_r", "ln_l", "ln_r", "ldesc_l", "ldesc_r")}
            p = port.stereo_points(d["kp_l"], d["pdesc_l"], d["kp_r"], d["pdesc_r"], 64.0 / w, 48.0 / h, cam,
                                   matching_s_ws=int(g("matching_s_ws", 10)), ratio=g("ratio", 0.9), best_lr=bool(g("best_lr", 1)),
                                   max_dist_epip=g("max_dist_epip", 1.0), min_disp=g("min_disp", 1.0))
            assert np.array_equal(p["kept_i1"], z[f"f{f}_c{c}_pt_kept_i1"])
            assert np.array_equal(bits(p["disp"]), bits(z[f"f{f}_c{c}_pt_disp"]))
            assert np.array_equal(bits(p["P"]), bits(z[f"f{f}_c{c}_pt_P"]))
            q = port.stereo_lines(d["ln_l"], d["ldesc_l"], d["ln_r"], d["ldesc_r"], 64.0 / w, 48.0 / h, cam,
                                  matching_s_ws=int(g("matching_s_ws", 10)), ratio=g("ratio", 0.9), best_lr=bool(g("best_lr", 1)),
                                  min_disp=g("min_disp", 1.0), line_horiz_th=g("line_horiz_th", 0.1),
                                  stereo_overlap_th=g("stereo_overlap_th", 0.75), ls_min_disp_ratio=g("ls_min_disp_ratio", 0.7))
            assert np.array_equal(q["kept_i1"], z[f"f{f}_c{c}_ls_kept_i1"])
            for key in ("disp_se", "sP", "eP", "le"):
                assert np.array_equal(bits(q[key]), bits(z[f"f{f}_c{c}_ls_{key}"])), key
    import ctypes as C
    lib = port.lib
    lib.plo_line_segment_overlap.restype = C.c_double
    lib.plo_line_overlap_stereo.restype = C.c_double
    lib.plo_line_overlap_stereo.argtypes = [C.c_double] * 5
    p2 = C.c_double * 2
    got = np.array([lib.plo_line_segment_overlap(p2(*r[0:2]), p2(*r[2:4]), p2(*r[4:6]), p2(*r[6:8])) for r in z["ov_in"]])
    assert np.array_equal(bits(got), bits(z["ov_out"]))
    got = np.array([lib.plo_line_overlap_stereo(*r, 0.1) for r in z["ovs_in"]])
    assert np.array_equal(bits(got), bits(z["ovs_out"]))
