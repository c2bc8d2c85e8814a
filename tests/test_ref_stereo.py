"""Pins the stereo-driver / gate restatements of oracle/plm_oracle.c on the REFERENCE'S OWN CODE:
stvo-pl/src/{stereoFrame,stereoFeatures,pinholeStereoCamera}.cpp compiled unmodified into
oracle/_ref/libplref_stereo.so (oracle/Makefile; stand-in OpenCV / Eigen headers under oracle/shim_stereo/).

Bit-exact comparison (fp64 outputs by their bit patterns) of
  StereoFrame::matchStereoPoints        stereoFrame.cpp:131-184   vs  plo_stereo_points
  StereoFrame::matchStereoLines         stereoFrame.cpp:320-409   vs  plo_stereo_lines
  filterLineSegmentDisparity            stereoFrame.cpp:416-426   vs  plo_stereo_filter_lines
  lineSegmentOverlapStereo              stereoFrame.cpp:484-519   vs  plo_line_overlap_stereo
  lineSegmentOverlap                    stereoFrame.cpp:521-627   vs  plo_line_segment_overlap
  PinholeStereoCamera::backProjection   pinholeStereoCamera.cpp:229-237
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from pl_inertial_slam_b200 import synth

port, rs = oracle.port, oracle.ref_stereo
pytestmark = pytest.mark.skipif(not rs.available(), reason="reference stereo build absent (needs /root/reference once)")

W, H = synth.IMG_W, synth.IMG_H
FX, FY, CX, CY, B = 435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297, 0.110073808127187
CAM_PORT = np.array([B, FX, CX, CY])          # oracle.port / plmatch.h order
CAM_REF = np.array([FX, FY, CX, CY, B])


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def same(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def degenerate_pair(seed):
    """A stereo pair salted with the degenerate inputs the gates special-case."""
    sp = synth.make_stereo_pair(synth.SEED0 + 200 + seed, n_pts=350, n_lines=180)
    rng = np.random.default_rng(seed)
    ln_l, ln_r = sp.ln_l.copy(), sp.ln_r.copy()
    ln_r[::11, 3] = ln_r[::11, 1]                       # horizontal right segment: 0/0 and x/0 in the interpolation
    ln_l[5::13, 3] = ln_l[5::13, 1]                     # horizontal left segment (lineHorizTh gate)
    ln_l[7::17, 2:] = ln_l[7::17, :2]                   # zero-length left segment: NaN line equation
    ln_r[3::19, 2:] = ln_r[3::19, :2]                   # zero-length right segment: NaN direction passes matchGrid
    ln_l[2::23, [0, 2]] = ln_l[2::23, [2, 0]]           # reversed endpoints
    kp_l, kp_r = sp.kp_l.copy(), sp.kp_r.copy()
    k = rng.choice(len(kp_l), 40, replace=False)
    kp_l[k, 0] = np.float32(W + 30.0)                   # off-grid left keypoints (window clipped to nothing)
    kp_r[::9, 1] += np.float32(1.0)                     # epipolar distance exactly around maxDistEpip
    return kp_l, sp.pdesc_l, kp_r, sp.pdesc_r, ln_l, sp.ldesc_l, ln_r, sp.ldesc_r


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("ratio,best_lr,ws", [(0.9, True, 10), (0.75, True, 10), (0.9, False, 10), (0.9, True, 3)])
def test_stereo_drivers_port_vs_reference(seed, ratio, best_lr, ws):
    kp_l, dpl, kp_r, dpr, ln_l, dll, ln_r, dlr = degenerate_pair(seed)
    p = port.stereo_points(kp_l, dpl, kp_r, dpr, synth.INV_W, synth.INV_H, CAM_PORT, matching_s_ws=ws, ratio=ratio, best_lr=best_lr)
    r = rs.stereo_points(kp_l, dpl, kp_r, dpr, W, H, CAM_REF, matching_s_ws=ws, ratio=ratio, best_lr=best_lr)
    assert len(r["kept_i1"]) > (100 if ws == 10 else 30)
    assert np.array_equal(p["kept_i1"], r["kept_i1"])
    assert same(p["disp"], r["disp"]) and same(p["P"], r["P"])
    assert np.array_equal(r["desc"], dpl[r["kept_i1"]])            # pdesc_l compaction (stereoFrame.cpp:172,183)
    assert same(r["pl"], kp_l[r["kept_i1"]].astype(np.float64))
    q = port.stereo_lines(ln_l, dll, ln_r, dlr, synth.INV_W, synth.INV_H, CAM_PORT, matching_s_ws=ws, ratio=ratio, best_lr=best_lr)
    s = rs.stereo_lines(ln_l, dll, ln_r, dlr, W, H, CAM_REF, matching_s_ws=ws, ratio=ratio, best_lr=best_lr)
    assert len(s["kept_i1"]) > (40 if ws == 10 else 10)
    assert np.array_equal(q["kept_i1"], s["kept_i1"])
    for key in ("disp_se", "sP", "eP", "le"):
        assert same(q[key], s[key]), key
    assert np.array_equal(s["desc"], dll[s["kept_i1"]])


def test_stereo_driver_thresholds_and_initial_flag():
    kp_l, dpl, kp_r, dpr, ln_l, dll, ln_r, dlr = degenerate_pair(9)
    for cfg in (dict(max_dist_epip=0.5, min_disp=20.0), dict(max_dist_epip=2.5, min_disp=0.0)):
        p = port.stereo_points(kp_l, dpl, kp_r, dpr, synth.INV_W, synth.INV_H, CAM_PORT, **cfg)
        r = rs.stereo_points(kp_l, dpl, kp_r, dpr, W, H, CAM_REF, **cfg)
        assert np.array_equal(p["kept_i1"], r["kept_i1"]) and same(p["disp"], r["disp"]) and same(p["P"], r["P"])
    for cfg in (dict(min_disp=15.0, line_horiz_th=2.0, stereo_overlap_th=0.9, ls_min_disp_ratio=0.9),
                dict(min_disp=0.0, line_horiz_th=0.0, stereo_overlap_th=0.1, ls_min_disp_ratio=0.2)):
        q = port.stereo_lines(ln_l, dll, ln_r, dlr, synth.INV_W, synth.INV_H, CAM_PORT, **cfg)
        s = rs.stereo_lines(ln_l, dll, ln_r, dlr, W, H, CAM_REF, **cfg)
        assert np.array_equal(q["kept_i1"], s["kept_i1"])
        for key in ("disp_se", "sP", "eP", "le"):
            assert same(q[key], s[key]), key
    # frame 0 numbers the stereo features 0..n-1, later frames mark them -1 (stereoFrame.cpp:175-178, :391-404);
    # level / sigma2 follow the octave (stereoFeatures.cpp:41-47)
    octv = np.arange(len(kp_l)) % 4
    a = rs.stereo_points(kp_l, dpl, kp_r, dpr, W, H, CAM_REF, initial=True, octave=octv)
    b = rs.stereo_points(kp_l, dpl, kp_r, dpr, W, H, CAM_REF, initial=False, octave=octv)
    assert (a["idx"] == np.arange(len(a["idx"]))).all() and (b["idx"] == -1).all()
    assert set(a["level"]) <= {0, 1, 2, 3} and same(a["P"], b["P"])
    s2 = np.array([1.0 / (1.2 ** int(l)) ** 2 for l in a["level"]])
    assert np.allclose(a["sigma2"], s2, rtol=1e-6)


def test_empty_and_tiny_frames():
    rng = np.random.default_rng(2)
    d = synth.rand_desc(rng, 3)
    kp = np.array([[100, 100], [200, 120], [300, 140]], np.float32)
    for n_l, n_r in ((0, 3), (3, 0), (1, 1), (3, 3)):
        r = rs.stereo_points(kp[:n_l], d[:n_l], kp[:n_r] - np.float32([5, 0]), d[:n_r], W, H, CAM_REF)
        p = port.stereo_points(kp[:n_l], d[:n_l], kp[:n_r] - np.float32([5, 0]), d[:n_r], synth.INV_W, synth.INV_H, CAM_PORT)
        assert np.array_equal(p["kept_i1"], r["kept_i1"]) and same(p["P"], r["P"])
        if n_l == 0 or n_r == 0:
            assert len(r["kept_i1"]) == 0
        elif n_l == n_r:
            assert len(r["kept_i1"]) == n_l      # identical descriptors 5 px apart on the same row: all kept


def test_scalar_gates_vs_reference():
    rng = np.random.default_rng(77)
    n = 4000
    # lineSegmentOverlapStereo: y-ranges incl. equal, nested, disjoint, near-horizontal observed segments
    v = rng.uniform(0, 480, (n, 4))
    v[::7, 1] = v[::7, 0] + rng.uniform(-0.2, 0.2, len(v[::7]))     # |dy| around lineHorizTh
    v[::11, 2:] = v[::11, :2]
    v[::13, 3] = v[::13, 2]                                          # zero-length projection
    lib = port.lib
    lib.plo_line_overlap_stereo.restype = C.c_double
    lib.plo_line_overlap_stereo.argtypes = [C.c_double] * 5
    got = np.array([lib.plo_line_overlap_stereo(*row, 0.1) for row in v])
    assert same(got, rs.line_overlap_stereo(v, 0.1))
    # lineSegmentOverlap: vertical (|dx| < 1), horizontal (|dy| < 1) and general observed segments
    s = rng.uniform(0, 700, (n, 8))
    s[::5, 2] = s[::5, 0] + rng.uniform(-1.5, 1.5, len(s[::5]))      # around the vertical special case
    s[1::5, 3] = s[1::5, 1] + rng.uniform(-1.5, 1.5, len(s[1::5]))   # around the horizontal special case
    s[2::31, 2:4] = s[2::31, 0:2]                                    # zero-length observed segment (0/0)
    lib.plo_line_segment_overlap.restype = C.c_double
    p2 = C.c_double * 2
    got = np.array([lib.plo_line_segment_overlap(p2(*r[0:2]), p2(*r[2:4]), p2(*r[4:6]), p2(*r[6:8])) for r in s])
    assert same(got, rs.line_overlap(s))
    # backProjection
    uvd = np.concatenate([rng.uniform(0, 752, (n, 2)), rng.uniform(0.5, 120, (n, 1))], 1)
    uvd[::50, 2] = 0.0                                               # zero disparity -> inf / NaN
    want = rs.back_projection(CAM_REF, uvd)
    with np.errstate(all="ignore"):
        bd = B / uvd[:, 2]
    mine = np.stack([bd * (uvd[:, 0] - CX), bd * (uvd[:, 1] - CY), bd * FX], 1)
    assert np.array_equal(bits(want), bits(mine))
    # filterLineSegmentDisparity (both overloads)
    spl, epl, spr, epr = (rng.uniform(0, 700, (n, 2)) for _ in range(4))
    d = rs.filter_line_disparity(spl, epl, spr, epr, 0.7)
    ds, de = spl[:, 0] - spr[:, 0], epl[:, 0] - epr[:, 0]
    with np.errstate(all="ignore"):
        bad = np.minimum(ds, de) / np.maximum(ds, de) < 0.7
    assert same(d[:, 0], np.where(bad, -1.0, ds)) and same(d[:, 1], np.where(bad, -1.0, de))
