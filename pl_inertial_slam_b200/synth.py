"""Synthetic EuRoC-shaped inputs for the matching path.

No dataset, vocabulary or detector is available offline, so every workload is generated:
752x480 stereo pairs with ~600 ORB-like and ~200 LBD-like 256-bit descriptors, a true match being
the base descriptor with Binomial(256, p) flipped bits, plus independent outliers.  All generators
are deterministic in (seed, sizes).  Seeds follow SURVEY 8d: SEED0 + config number.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import grid as G

SEED0 = 20261018
IMG_W, IMG_H = 752, 480
INV_W = G.GRID_COLS / float(IMG_W)   # stereoFrame.cpp:47
INV_H = G.GRID_ROWS / float(IMG_H)   # stereoFrame.cpp:48


def rand_desc(rng: np.random.Generator, n: int) -> np.ndarray:
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def flip_bits(rng: np.random.Generator, desc: np.ndarray, p: float = 0.08) -> np.ndarray:
    """Each of the 256 bits flips independently with probability p."""
    noise = rng.random((desc.shape[0], 256)) < p
    return desc ^ np.packbits(noise, axis=1)


def tie_stress_desc(rng: np.random.Generator, n: int, n_dup: Optional[int] = None) -> np.ndarray:
    """Descriptors with only 4 non-zero bytes and many exact duplicates: forces equal distances so
    that lowest-index tie-breaking is exercised (SURVEY 8d 'tie-stress set')."""
    d = np.zeros((n, 32), np.uint8)
    d[:, :4] = rng.integers(0, 4, size=(n, 4), dtype=np.uint8)
    n_dup = n // 3 if n_dup is None else n_dup
    if n > 1 and n_dup > 0:
        src = rng.integers(0, n, size=n_dup)
        dst = rng.integers(0, n, size=n_dup)
        d[dst] = d[src]
    return d


@dataclass
class StereoPair:
    kp_l: np.ndarray      # n x 2 float32 pixel coords (cv::KeyPoint::pt)
    kp_r: np.ndarray
    pdesc_l: np.ndarray   # n x 32 uint8
    pdesc_r: np.ndarray
    ln_l: np.ndarray      # m x 4 float32 (sx, sy, ex, ey)  (cv::line_descriptor::KeyLine)
    ln_r: np.ndarray
    ldesc_l: np.ndarray
    ldesc_r: np.ndarray


def make_stereo_pair(seed: int, n_pts: int = 600, n_lines: int = 200, outlier_frac: float = 0.25,
                     flip_p: float = 0.08) -> StereoPair:
    """Config 1: right = left shifted by a disparity U[1,100] px, y jitter N(0, 0.5), 25 % of the
    right features are unrelated outliers, order of the right features is shuffled."""
    rng = np.random.default_rng(seed)
    b = 19.0  # orbEdgeTh border (config.cpp:96)
    xl = rng.uniform(b, IMG_W - b, n_pts)
    yl = rng.uniform(b, IMG_H - b, n_pts)
    disp = rng.uniform(1.0, 100.0, n_pts)
    xr = xl - disp
    yr = yl + rng.normal(0.0, 0.5, n_pts)
    pdesc_l = rand_desc(rng, n_pts)
    pdesc_r = flip_bits(rng, pdesc_l, flip_p)
    out = rng.random(n_pts) < outlier_frac
    n_out = int(out.sum())
    xr[out] = rng.uniform(b, IMG_W - b, n_out)
    yr[out] = rng.uniform(b, IMG_H - b, n_out)
    pdesc_r[out] = rand_desc(rng, n_out)
    perm = rng.permutation(n_pts)
    kp_l = np.stack([xl, yl], 1).astype(np.float32)
    kp_r = np.stack([xr, yr], 1)[perm].astype(np.float32)
    pdesc_r = np.ascontiguousarray(pdesc_r[perm])

    min_len = 0.025 * IMG_H  # minLineLength relative to the image (config.cpp:66)
    sx = rng.uniform(b, IMG_W - b, n_lines)
    sy = rng.uniform(b, IMG_H - b, n_lines)
    ang = rng.uniform(0, 2 * np.pi, n_lines)
    length = rng.uniform(min_len, 150.0, n_lines)
    ex = np.clip(sx + length * np.cos(ang), 1.0, IMG_W - 2.0)
    ey = np.clip(sy + length * np.sin(ang), 1.0, IMG_H - 2.0)
    ldisp = rng.uniform(1.0, 100.0, n_lines)
    sxr = sx - ldisp + rng.normal(0, 0.3, n_lines)
    exr = ex - ldisp + rng.normal(0, 0.3, n_lines)
    syr = sy + rng.normal(0, 0.5, n_lines)
    eyr = ey + rng.normal(0, 0.5, n_lines)
    ldesc_l = rand_desc(rng, n_lines)
    ldesc_r = flip_bits(rng, ldesc_l, flip_p)
    lout = rng.random(n_lines) < outlier_frac
    n_lo = int(lout.sum())
    sxr[lout] = rng.uniform(b, IMG_W - b, n_lo); syr[lout] = rng.uniform(b, IMG_H - b, n_lo)
    exr[lout] = rng.uniform(b, IMG_W - b, n_lo); eyr[lout] = rng.uniform(b, IMG_H - b, n_lo)
    ldesc_r[lout] = rand_desc(rng, n_lo)
    lperm = rng.permutation(n_lines)
    ln_l = np.stack([sx, sy, ex, ey], 1).astype(np.float32)
    ln_r = np.stack([sxr, syr, exr, eyr], 1)[lperm].astype(np.float32)
    ldesc_r = np.ascontiguousarray(ldesc_r[lperm])
    return StereoPair(kp_l, kp_r, pdesc_l, pdesc_r, ln_l, ln_r, ldesc_l, ldesc_r)


# ---- host-side preparation of the matchGrid arguments, as the reference's drivers do it ---------

def stereo_points_grid_args(sp: StereoPair, matching_s_ws: int = 10) -> Dict:
    """Arguments of the matchGrid call at stereoFrame.cpp:157: coords = trunc(kp_l * inv), grid over
    the right keypoints, window = (matchingSWs, 0) x (0, 0)."""
    xy = np.stack([np.trunc(sp.kp_l[:, 0].astype(np.float64) * INV_W),
                   np.trunc(sp.kp_l[:, 1].astype(np.float64) * INV_H)], 1).astype(np.int32)
    cs, ci = G.csr_from_points(sp.kp_r[:, 0].astype(np.float64) * INV_W,
                               sp.kp_r[:, 1].astype(np.float64) * INV_H)
    return dict(xy=xy, d1=sp.pdesc_l, cell_start=cs, cell_items=ci, rows=G.GRID_ROWS, cols=G.GRID_COLS,
                d2=sp.pdesc_r, win=np.array([matching_s_ws, 0, 0, 0], np.int32))


def stereo_lines_grid_args(sp: StereoPair, matching_s_ws: int = 10) -> Dict:
    """Arguments of the matchGrid call at stereoFrame.cpp:356 (coords :329-333, grid + directions
    :336-349)."""
    l = sp.ln_l.astype(np.float64)
    r = sp.ln_r.astype(np.float64)
    xyxy = np.stack([np.trunc(l[:, 0] * INV_W), np.trunc(l[:, 1] * INV_H),
                     np.trunc(l[:, 2] * INV_W), np.trunc(l[:, 3] * INV_H)], 1).astype(np.int32)
    cs, ci = G.csr_from_lines(r[:, 0] * INV_W, r[:, 1] * INV_H, r[:, 2] * INV_W, r[:, 3] * INV_H)
    fr = sp.ln_r
    vx = (fr[:, 2] - fr[:, 0]).astype(np.float64) * INV_W   # float subtraction, then * double
    vy = (fr[:, 3] - fr[:, 1]).astype(np.float64) * INV_H
    with np.errstate(invalid="ignore", divide="ignore"):
        mag = np.sqrt(vx * vx + vy * vy)
        dirs2 = np.stack([vx / mag, vy / mag], 1)
    return dict(xyxy=xyxy, d1=sp.ldesc_l, cell_start=cs, cell_items=ci, rows=G.GRID_ROWS, cols=G.GRID_COLS,
                d2=sp.ldesc_r, dirs2=dirs2, win=np.array([matching_s_ws, 0, 0, 0], np.int32))


def make_temporal_pair(seed: int, n_pts: int = 600, n_lines: int = 200, flip_p: float = 0.08,
                       outlier_frac: float = 0.25):
    """Config 2: prev/curr left frames; curr = prev moved by U[-30, 30] px.  Returns two StereoPair-like
    objects sharing structure (only the left halves are used by the temporal matchers)."""
    prev = make_stereo_pair(seed, n_pts, n_lines, outlier_frac, flip_p)
    rng = np.random.default_rng(seed + 7919)
    kp = prev.kp_l.astype(np.float64) + rng.uniform(-30, 30, (n_pts, 2))
    pdesc = flip_bits(rng, prev.pdesc_l, flip_p)
    out = rng.random(n_pts) < outlier_frac
    pdesc[out] = rand_desc(rng, int(out.sum()))
    perm = rng.permutation(n_pts)
    shift = rng.uniform(-30, 30, (n_lines, 2))
    ln = prev.ln_l.astype(np.float64) + np.concatenate([shift, shift], 1)
    ldesc = flip_bits(rng, prev.ldesc_l, flip_p)
    lout = rng.random(n_lines) < outlier_frac
    ldesc[lout] = rand_desc(rng, int(lout.sum()))
    lperm = rng.permutation(n_lines)
    curr = StereoPair(kp[perm].astype(np.float32), prev.kp_r, np.ascontiguousarray(pdesc[perm]), prev.pdesc_r,
                      ln[lperm].astype(np.float32), prev.ln_r, np.ascontiguousarray(ldesc[lperm]), prev.ldesc_r)
    return prev, curr


def kf_points_grid_args(prev: StereoPair, curr: StereoPair, ws: int = 3) -> Dict:
    """matchKF2KFPoints-shaped call (mapHandler.cpp:300-322): projected prev points as queries, grid
    over curr points, window +-matchingF2FWs in both axes."""
    p = prev.kp_l.astype(np.float64)
    xy = np.stack([np.trunc(p[:, 0] * INV_W), np.trunc(p[:, 1] * INV_H)], 1).astype(np.int32)
    c = curr.kp_l.astype(np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * INV_W, c[:, 1] * INV_H)
    return dict(xy=xy, d1=prev.pdesc_l, cell_start=cs, cell_items=ci, rows=G.GRID_ROWS, cols=G.GRID_COLS,
                d2=curr.pdesc_l, win=np.array([ws, ws, ws, ws], np.int32))


def kf_lines_grid_args(prev: StereoPair, curr: StereoPair, ws: int = 3, pixel_coord_quirk: bool = False) -> Dict:
    """matchKF2KFLines-shaped call (mapHandler.cpp:431-469).  With pixel_coord_quirk the query
    coordinates are raw pixels, not grid cells, as in the reference (:443-444)."""
    l = prev.ln_l.astype(np.float64)
    sx, sy = (1.0, 1.0) if pixel_coord_quirk else (INV_W, INV_H)
    xyxy = np.stack([np.trunc(l[:, 0] * sx), np.trunc(l[:, 1] * sy),
                     np.trunc(l[:, 2] * sx), np.trunc(l[:, 3] * sy)], 1).astype(np.int32)
    r = curr.ln_l.astype(np.float64)
    cs, ci = G.csr_from_lines(r[:, 0] * INV_W, r[:, 1] * INV_H, r[:, 2] * INV_W, r[:, 3] * INV_H)
    vx = (r[:, 2] - r[:, 0]) * INV_W
    vy = (r[:, 3] - r[:, 1]) * INV_H
    with np.errstate(invalid="ignore", divide="ignore"):
        mag = np.sqrt(vx * vx + vy * vy)
        dirs2 = np.stack([vx / mag, vy / mag], 1)
    return dict(xyxy=xyxy, d1=prev.ldesc_l, cell_start=cs, cell_items=ci, rows=G.GRID_ROWS, cols=G.GRID_COLS,
                d2=curr.ldesc_l, dirs2=dirs2, win=np.array([ws, ws, ws, ws], np.int32))


def make_map_points(seed: int, n_map: int, frame: StereoPair, frac_seen: float = 0.002):
    """Config 4: a local map of n_map representative descriptors with projected cell coordinates
    uniform over the grid (some just off-grid); a small fraction are noisy copies of frame features
    so that true matches exist."""
    rng = np.random.default_rng(seed)
    desc = rand_desc(rng, n_map)
    xy = np.stack([rng.integers(-2, G.GRID_COLS + 2, n_map), rng.integers(-2, G.GRID_ROWS + 2, n_map)], 1)
    n_f = frame.pdesc_l.shape[0]
    n_seen = max(1, int(n_map * frac_seen))
    rows = rng.choice(n_map, n_seen, replace=False)
    src = rng.integers(0, n_f, n_seen)
    desc[rows] = flip_bits(rng, frame.pdesc_l[src], 0.08)
    f = frame.kp_l.astype(np.float64)
    xy[rows, 0] = np.trunc(f[src, 0] * INV_W) + rng.integers(-2, 3, n_seen)
    xy[rows, 1] = np.trunc(f[src, 1] * INV_H) + rng.integers(-2, 3, n_seen)
    return np.ascontiguousarray(desc), xy.astype(np.int32)


def make_keyframe_db(seed: int, n_kf: int, per_kf: int = 800, revisit_frac: float = 0.05, flip_p: float = 0.08,
                     chunk: int = 256) -> np.ndarray:
    """Config 5: keyframe descriptor database, n_kf x per_kf rows; revisit_frac of the keyframes are
    noisy copies of an earlier keyframe (loop revisits)."""
    rng = np.random.default_rng(seed)
    db = np.empty((n_kf * per_kf, 32), np.uint8)
    for k0 in range(0, n_kf, chunk):
        k1 = min(n_kf, k0 + chunk)
        db[k0 * per_kf:k1 * per_kf] = rng.integers(0, 256, size=((k1 - k0) * per_kf, 32), dtype=np.uint8)
    n_rev = int(n_kf * revisit_frac)
    if n_rev and n_kf > 1:
        dst = rng.choice(np.arange(1, n_kf), size=min(n_rev, n_kf - 1), replace=False)
        for k in dst:
            src = int(rng.integers(0, k))
            db[k * per_kf:(k + 1) * per_kf] = flip_bits(rng, db[src * per_kf:(src + 1) * per_kf], flip_p)
    return db
