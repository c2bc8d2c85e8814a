"""Host-side mirrors of the reference's MATCHING DRIVERS: the functions that build the arguments of
match / matchGrid, call them and apply the per-match gates.  Only the
matching half of each driver is mirrored; building PointFeature / LineFeature objects, back-projection
and the pose maths stay with the caller.  Every arithmetic step runs on the GPU through matching.py.

    matchStereoPoints / matchStereoLines     stvo-pl/src/stereoFrame.cpp:131-184 / :320-409
    matchF2FPoints / matchF2FLines           stvo-pl/src/stereoFrameHandler.cpp:158-207
    matchKF2KFPoints / matchKF2KFLines       src/mapHandler.cpp:285-330 / :416-478
    matchMap2KFPoints / matchMap2KFLines     src/mapHandler.cpp:583-650 / :685-765   (same call shape)
    loopClosureMatch                         src/mapHandler.cpp:3325-3400 (isLoopClosure)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from . import grid as G
from . import matching as M


class SlamConfig:
    """The SlamConfig values the drivers read (src/slamConfig.cpp:36-87 defaults, EuRoC yaml in brackets)."""
    fastMatching: bool = True        # slamConfig.cpp:43 false [config_euroc.yaml: true]
    minPointMatches: int = 10
    minLineMatches: int = 6
    lcInlierRatio: float = 30.0
    maxKFEpipP: float = 1.0          # slamConfig.cpp:51
    maxKFEpipL: float = 1.0          # slamConfig.cpp:52


@dataclass
class StereoMatches:
    matches_12: np.ndarray     # per left feature: right index or -1 (after matchGrid)
    keep: np.ndarray           # uint8 per left feature: survived the stereo gates
    disparity: np.ndarray      # points: n; lines: n x 2 (disp_s, disp_e)
    desc_kept: np.ndarray      # the rows of the left descriptors that survive (pdesc_l_ = pdesc_l_aux, :183/:408)


def _scaled_cells(xy_px: np.ndarray, inv_w: float, inv_h: float) -> np.ndarray:
    """std::make_pair(px * inv_width, py * inv_height) narrowed to pair<int,int>: truncation toward 0."""
    p = np.asarray(xy_px, np.float64)
    return np.stack([np.trunc(p[:, 0] * inv_w), np.trunc(p[:, 1] * inv_h)], 1).astype(np.int32)


def matchStereoPoints(points_l: np.ndarray, points_r: np.ndarray, pdesc_l: np.ndarray, pdesc_r: np.ndarray,
                      inv_width: float, inv_height: float, ctx=None) -> Optional[StereoMatches]:
    """points_*: n x 2 float32 pixel coordinates (cv::KeyPoint::pt)."""
    if len(points_l) == 0 or len(points_r) == 0:      # stereoFrame.cpp:137-138
        return None
    pl = np.asarray(points_l, np.float32)
    pr = np.asarray(points_r, np.float32)
    coords = _scaled_cells(pl, inv_width, inv_height)
    cs, ci = G.csr_from_points(pr[:, 0].astype(np.float64) * inv_width, pr[:, 1].astype(np.float64) * inv_height)
    w = G.GridWindow((M.Config.matchingSWs, 0), (0, 0))
    m12: List[int] = []
    M.matchGrid(coords, pdesc_l, (cs, ci, G.GRID_ROWS, G.GRID_COLS), pdesc_r, w, m12, ctx=ctx)
    m = np.asarray(m12, np.int32)
    _, keep, disp = M.stereo_filter_points(pl, pr, m, ctx=ctx)
    return StereoMatches(m, keep, disp, np.ascontiguousarray(pdesc_l[keep.astype(bool)]))


def line_grid(lines_px: np.ndarray, inv_width: float, inv_height: float):
    """Grid + directions over line segments given in pixels (stereoFrame.cpp:336-349)."""
    r = np.asarray(lines_px, np.float32)
    r64 = r.astype(np.float64)
    cs, ci = G.csr_from_lines(r64[:, 0] * inv_width, r64[:, 1] * inv_height, r64[:, 2] * inv_width, r64[:, 3] * inv_height)
    vx = (r[:, 2] - r[:, 0]).astype(np.float64) * inv_width       # float subtraction, then * double (:343)
    vy = (r[:, 3] - r[:, 1]).astype(np.float64) * inv_height
    with np.errstate(invalid="ignore", divide="ignore"):
        mag = np.sqrt(vx * vx + vy * vy)
        dirs = np.stack([vx / mag, vy / mag], 1)
    return cs, ci, dirs


def matchStereoLines(lines_l: np.ndarray, lines_r: np.ndarray, ldesc_l: np.ndarray, ldesc_r: np.ndarray,
                     inv_width: float, inv_height: float, ctx=None) -> Optional[StereoMatches]:
    """lines_*: n x 4 float32 (startPointX, startPointY, endPointX, endPointY)."""
    if len(lines_l) == 0 or len(lines_r) == 0:
        return None
    ll = np.asarray(lines_l, np.float32)
    lr = np.asarray(lines_r, np.float32)
    l64 = ll.astype(np.float64)
    coords = np.stack([np.trunc(l64[:, 0] * inv_width), np.trunc(l64[:, 1] * inv_height),
                       np.trunc(l64[:, 2] * inv_width), np.trunc(l64[:, 3] * inv_height)], 1).astype(np.int32)
    cs, ci, dirs = line_grid(lr, inv_width, inv_height)
    w = G.GridWindow((M.Config.matchingSWs, 0), (0, 0))
    m12: List[int] = []
    M.matchGrid(coords, ldesc_l, (cs, ci, G.GRID_ROWS, G.GRID_COLS), ldesc_r, dirs, w, m12, ctx=ctx)
    m = np.asarray(m12, np.int32)
    _, keep, disp = M.stereo_filter_lines(ll, lr, m, ctx=ctx)
    return StereoMatches(m, keep, disp, np.ascontiguousarray(ldesc_l[keep.astype(bool)]))


def matchF2FPoints(prev_pdesc_l: np.ndarray, curr_pdesc_l: np.ndarray, ctx=None) -> np.ndarray:
    """stereoFrameHandler.cpp:158-180: match(prev, curr, minRatio12P); empty inputs -> no matches."""
    if len(prev_pdesc_l) == 0 or len(curr_pdesc_l) == 0:
        return np.zeros(0, np.int32)
    m12: List[int] = []
    M.match(prev_pdesc_l, curr_pdesc_l, np.float32(M.Config.minRatio12P), m12, ctx=ctx)
    return np.asarray(m12, np.int32)


def matchF2FLines(prev_ldesc_l: np.ndarray, curr_ldesc_l: np.ndarray, ctx=None, prev_lines_px: np.ndarray = None,
                  curr_lines_px: np.ndarray = None, overlap_th: float = None) -> np.ndarray:
    """stereoFrameHandler.cpp:182-207: match(prev, curr, minRatio12L).

    The fork applies no geometric filter here.  Passing the two segment arrays (n x (sx, sy, ex, ey) pixels) and
    ``overlap_th`` turns on the opt-in overlap / angle filter of BASELINE config 2 (M.line_pair_filter:
    StereoFrame::lineSegmentOverlap + the lineSimTh direction test); filtered matches are set to -1."""
    if len(prev_ldesc_l) == 0 or len(curr_ldesc_l) == 0:
        return np.zeros(0, np.int32)
    m12: List[int] = []
    M.match(prev_ldesc_l, curr_ldesc_l, np.float32(M.Config.minRatio12L), m12, ctx=ctx)
    out = np.asarray(m12, np.int32)
    if overlap_th is not None and prev_lines_px is not None and curr_lines_px is not None:
        _, keep, _, _ = M.line_pair_filter(prev_lines_px, curr_lines_px, out, overlap_th, ctx=ctx)
        out = np.where(keep.astype(bool), out, -1).astype(np.int32)
    return out


def _grid_then_fallback(coords, desc1, grid, desc2, dirs2, ws, min_matches, nnr, n1_feats, n2_feats, ctx):
    """Shared shape of matchKF2KF* / matchMap2KF*: matchGrid with a +-ws window when fastMatching,
    then match() ON THE SAME VECTOR when too few matches were found (stale entries survive)."""
    matches = 0
    m12: List[int] = []
    if SlamConfig.fastMatching:
        w = G.GridWindow((ws, ws), (ws, ws))
        if dirs2 is None:
            matches = M.matchGrid(coords, desc1, grid, desc2, w, m12, ctx=ctx)
        else:
            matches = M.matchGrid(coords, desc1, grid, desc2, dirs2, w, m12, ctx=ctx)
    if n2_feats > min_matches and n1_feats > min_matches and matches < min_matches:
        matches = M.match(desc1, desc2, np.float32(nnr), m12, ctx=ctx)
    return matches, np.asarray(m12, np.int32)


def matchKF2KFPoints(pj_points: np.ndarray, prev_pdesc_l: np.ndarray, curr_points_px: np.ndarray,
                     curr_pdesc_l: np.ndarray, inv_width: float, inv_height: float, ctx=None):
    """mapHandler.cpp:285-330.  pj_points: projections of the previous keyframe's 3-D points into the
    current image (pixels); curr_points_px: the current keyframe's stereo points (pixels)."""
    if len(prev_pdesc_l) == 0 or len(curr_pdesc_l) == 0:
        return 0, np.zeros(0, np.int32)
    coords = _scaled_cells(pj_points, inv_width, inv_height)
    c = np.asarray(curr_points_px, np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * inv_width, c[:, 1] * inv_height)
    return _grid_then_fallback(coords, prev_pdesc_l, (cs, ci, G.GRID_ROWS, G.GRID_COLS), curr_pdesc_l, None,
                               M.Config.matchingF2FWs, SlamConfig.minPointMatches, M.Config.minRatio12P,
                               len(prev_pdesc_l), len(curr_pdesc_l), ctx)


def matchKF2KFLines(pj_lines_px: np.ndarray, prev_ldesc_l: np.ndarray, curr_lines_px: np.ndarray,
                    curr_ldesc_l: np.ndarray, inv_width: float, inv_height: float, ctx=None):
    """mapHandler.cpp:416-478.  NOTE the reference passes the projected query lines in PIXELS, not grid
    cells (:443-444) -- kept."""
    if len(prev_ldesc_l) == 0 or len(curr_ldesc_l) == 0:
        return 0, np.zeros(0, np.int32)
    coords = np.trunc(np.asarray(pj_lines_px, np.float64)).astype(np.int32).reshape(-1, 4)
    cs, ci, dirs = line_grid(curr_lines_px, inv_width, inv_height)
    return _grid_then_fallback(coords, prev_ldesc_l, (cs, ci, G.GRID_ROWS, G.GRID_COLS), curr_ldesc_l, dirs,
                               M.Config.matchingF2FWs, SlamConfig.minLineMatches, M.Config.minRatio12L,
                               len(prev_ldesc_l), len(curr_ldesc_l), ctx)


def matchMap2KFPoints(pj_points: np.ndarray, map_lpt_desc: np.ndarray, unmatched_points_px: np.ndarray,
                      unmatched_pt_desc: np.ndarray, inv_width: float, inv_height: float, ctx=None):
    """mapHandler.cpp:583-650: local-map representative descriptors (projected into the image) against
    the keyframe's still unmatched points."""
    if len(map_lpt_desc) == 0 or len(unmatched_pt_desc) == 0:
        return 0, np.zeros(0, np.int32)
    coords = _scaled_cells(pj_points, inv_width, inv_height)
    c = np.asarray(unmatched_points_px, np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * inv_width, c[:, 1] * inv_height)
    return _grid_then_fallback(coords, map_lpt_desc, (cs, ci, G.GRID_ROWS, G.GRID_COLS), unmatched_pt_desc, None,
                               M.Config.matchingF2FWs, SlamConfig.minPointMatches, M.Config.minRatio12P,
                               len(map_lpt_desc), len(coords), ctx)


def matchMap2KFLines(pj_lines_px: np.ndarray, map_lls_desc: np.ndarray, unmatched_lines_px: np.ndarray,
                     unmatched_ls_desc: np.ndarray, inv_width: float, inv_height: float, ctx=None):
    """mapHandler.cpp:685-765 (query lines scaled to grid cells here, :733-736)."""
    if len(map_lls_desc) == 0 or len(unmatched_ls_desc) == 0:
        return 0, np.zeros(0, np.int32)
    l = np.asarray(pj_lines_px, np.float64).reshape(-1, 4)
    coords = np.stack([np.trunc(l[:, 0] * inv_width), np.trunc(l[:, 1] * inv_height),
                       np.trunc(l[:, 2] * inv_width), np.trunc(l[:, 3] * inv_height)], 1).astype(np.int32)
    cs, ci, dirs = line_grid(unmatched_lines_px, inv_width, inv_height)
    return _grid_then_fallback(coords, map_lls_desc, (cs, ci, G.GRID_ROWS, G.GRID_COLS), unmatched_ls_desc, dirs,
                               M.Config.matchingF2FWs, SlamConfig.minLineMatches, M.Config.minRatio12L,
                               len(map_lls_desc), len(coords), ctx)


def map_view(Twf: np.ndarray, cam: Sequence[float], inv_width: float, inv_height: float, width: int, height: int) -> L.MapView:
    """Twf (4x4 or 3x4, world -> frame), cam = (fx, fy, cx, cy)."""
    v = L.MapView()
    T = np.asarray(Twf, np.float64).reshape(-1, 4)[:3]
    v.T[:] = [float(x) for x in T.reshape(-1)]
    v.fx, v.fy, v.cx, v.cy = [float(x) for x in cam]
    v.inv_width, v.inv_height, v.width, v.height = float(inv_width), float(inv_height), int(width), int(height)
    return v


def mapSelect(X: np.ndarray, active: Optional[np.ndarray], view: L.MapView, ctx=None):
    """The local-map selection of matchMap2KFPoints / Lines (mapHandler.cpp:596-609, :698-714): landmarks (n x 3
    points or n x 6 segments) that are active and project inside the image -> (sel, coords, pf), compacted in order."""
    X = np.ascontiguousarray(X, np.float64)
    is_lines = X.shape[1] == 6
    n, per = len(X), 2 if is_lines else 1
    act = None if active is None else np.ascontiguousarray(active, np.uint8)
    sel, coords, pf = np.zeros(max(n, 1), np.int32), np.zeros((max(n, 1), 2 * per), np.int32), np.zeros((max(n, 1), 2 * per))
    m = C.c_int(0)
    L.check(L.load().plm_map_select(M._h(ctx), int(is_lines), X.ctypes.data_as(L.f64p), None if act is None else act.ctypes.data_as(L.u8p),
                                    n, C.byref(view), sel.ctypes.data_as(L.i32p), coords.ctypes.data_as(L.i32p), pf.ctypes.data_as(L.f64p),
                                    C.byref(m)), "plm_map_select")
    return sel[:m.value].copy(), coords[:m.value].copy(), pf[:m.value].copy()


def mapGate(pf: np.ndarray, m12: np.ndarray, feat: np.ndarray, max_epip: float, count: int, ctx=None):
    """The reprojection gate after the matcher (mapHandler.cpp:652-680 points: feat = pl of the keyframe points;
    :767-800 lines: feat = line equations le) -> (count after the rejections, ok flags per compacted row)."""
    pf = np.ascontiguousarray(pf, np.float64)
    is_lines = pf.shape[1] == 4
    m = np.ascontiguousarray(m12, np.int32)
    feat = np.ascontiguousarray(feat, np.float64)
    ok = np.zeros(max(len(m), 1), np.uint8)
    c = C.c_int(int(count))
    L.check(L.load().plm_map_gate(M._h(ctx), int(is_lines), pf.ctypes.data_as(L.f64p), m.ctypes.data_as(L.i32p), len(m),
                                  feat.ctypes.data_as(L.f64p), len(feat), float(max_epip), ok.ctypes.data_as(L.u8p), C.byref(c)),
            "plm_map_gate")
    return c.value, ok[:len(m)].copy()


def matchMap2KFPointsFull(X: np.ndarray, active, med_desc: np.ndarray, view: L.MapView, unmatched_points_px: np.ndarray,
                          unmatched_pt_desc: np.ndarray, ctx=None):
    """MapHandler::matchMap2KFPoints end to end (mapHandler.cpp:583-682): selection, matchGrid (+ match fallback),
    epipolar gate -> dict(sel, m12, ok, matches)."""
    sel, coords, pf = mapSelect(X, active, view, ctx)
    if len(sel) == 0 or len(unmatched_pt_desc) == 0:
        return dict(sel=sel, m12=np.zeros(len(sel), np.int32) - 1, ok=np.zeros(len(sel), np.uint8), matches=0)
    c = np.asarray(unmatched_points_px, np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * view.inv_width, c[:, 1] * view.inv_height)
    d1 = np.ascontiguousarray(med_desc[sel])
    n, m12 = _grid_then_fallback(coords, d1, (cs, ci, G.GRID_ROWS, G.GRID_COLS), unmatched_pt_desc, None, M.Config.matchingF2FWs,
                                 SlamConfig.minPointMatches, M.Config.minRatio12P, len(d1), len(coords), ctx)
    n, ok = mapGate(pf, m12, c, SlamConfig.maxKFEpipP, n, ctx)
    return dict(sel=sel, m12=m12, ok=ok, matches=n)


def matchMap2KFLinesFull(X: np.ndarray, active, med_desc: np.ndarray, view: L.MapView, unmatched_lines_px: np.ndarray,
                         unmatched_le: np.ndarray, unmatched_ls_desc: np.ndarray, ctx=None):
    """MapHandler::matchMap2KFLines end to end (mapHandler.cpp:685-803)."""
    sel, coords, pf = mapSelect(X, active, view, ctx)
    if len(sel) == 0 or len(unmatched_ls_desc) == 0:
        return dict(sel=sel, m12=np.zeros(len(sel), np.int32) - 1, ok=np.zeros(len(sel), np.uint8), matches=0)
    cs, ci, dirs = line_grid(unmatched_lines_px, view.inv_width, view.inv_height)
    d1 = np.ascontiguousarray(med_desc[sel])
    n, m12 = _grid_then_fallback(coords, d1, (cs, ci, G.GRID_ROWS, G.GRID_COLS), unmatched_ls_desc, dirs, M.Config.matchingF2FWs,
                                 SlamConfig.minLineMatches, M.Config.minRatio12L, len(d1), len(coords), ctx)
    n, ok = mapGate(pf, m12, unmatched_le, SlamConfig.maxKFEpipL, n, ctx)
    return dict(sel=sel, m12=m12, ok=ok, matches=n)


def loopClosureMatch(kf0_pdesc: np.ndarray, kf1_pdesc: np.ndarray, kf0_ldesc: np.ndarray, kf1_ldesc: np.ndarray, ctx=None):
    """isLoopClosure's matching + inlier-ratio gate (mapHandler.cpp:3325-3400) -> dict."""
    common_pt, m_pt = 0, np.zeros(0, np.int32)
    if len(kf0_pdesc) and len(kf1_pdesc):
        v: List[int] = []
        common_pt = M.match(kf0_pdesc, kf1_pdesc, np.float32(M.Config.minRatio12P), v, ctx=ctx)
        m_pt = np.asarray(v, np.int32)
    common_ls, m_ls = 0, np.zeros(0, np.int32)
    if len(kf0_ldesc) and len(kf1_ldesc):
        v = []
        common_ls = M.match(kf0_ldesc, kf1_ldesc, np.float32(M.Config.minRatio12L), v, ctx=ctx)
        m_ls = np.asarray(v, np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_pt = max(np.float64(100.0 * common_pt) / len(kf0_pdesc), np.float64(100.0 * common_pt) / len(kf1_pdesc)) if len(kf0_pdesc) and len(kf1_pdesc) else 0.0
        r_ls = max(np.float64(100.0 * common_ls) / len(kf0_ldesc), np.float64(100.0 * common_ls) / len(kf1_ldesc)) if len(kf0_ldesc) and len(kf1_ldesc) else 0.0
    ok = bool(r_pt > SlamConfig.lcInlierRatio and r_ls > SlamConfig.lcInlierRatio)
    return dict(common_pt=common_pt, common_ls=common_ls, matches_pt=m_pt, matches_ls=m_ls, inl_ratio_pt=float(r_pt),
                inl_ratio_ls=float(r_ls), inl_ratio_condition=ok)
