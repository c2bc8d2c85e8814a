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


# ---- local-map landmarks: observation lists behind the med_desc rows (src/mapFeatures.cpp) -------

def make_landmark_observations(seed: int, n_lm: int, mean_obs: float = 8.0, max_obs: int = 32, flip_p: float = 0.06,
                               long_lists: int = 0, long_len: int = 100, tie: bool = False, empty_frac: float = 0.0):
    """Observation arenas of n_lm map landmarks: each landmark has a base descriptor and 1 + Poisson(mean_obs-1)
    observations (clipped to max_obs) that are noisy copies of it, plus unit observation directions.
    `long_lists` landmarks get `long_len` observations (the CTA kernel's path); `tie` uses the tie-stress
    descriptors (many equal medians -> first-row rule); `empty_frac` of the landmarks get no observation.
    Returns (desc_obs uint8[n_obs, 32], dir_obs float64[n_obs, 3], obs_start int32[n_lm + 1])."""
    rng = np.random.default_rng(seed)
    counts = np.clip(1 + rng.poisson(max(mean_obs - 1.0, 0.0), n_lm), 1, max_obs).astype(np.int64)
    if long_lists:
        counts[rng.choice(n_lm, min(long_lists, n_lm), replace=False)] = long_len
    if empty_frac > 0:
        counts[rng.random(n_lm) < empty_frac] = 0
    obs_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n_obs = int(obs_start[-1])
    owner = np.repeat(np.arange(n_lm), counts)
    if tie:
        desc = tie_stress_desc(rng, n_obs)
    else:
        desc = flip_bits(rng, rand_desc(rng, n_lm)[owner], flip_p) if n_obs else np.zeros((0, 32), np.uint8)
    dirs = rng.normal(size=(n_obs, 3))
    dirs /= np.maximum(np.linalg.norm(dirs, axis=1, keepdims=True), 1e-12)
    return np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(dirs, np.float64), obs_start


# ---- bag-of-words vocabulary (the DBoW2 blobs are not in the mount) -------------------------------

@dataclass
class FlatVoc:
    """A DBoW2 vocabulary tree as the flat arrays plm_voc_create takes (same attribute names as
    oracle.FlatVocabulary)."""
    child_start: np.ndarray
    child_ids: np.ndarray
    node_desc: np.ndarray
    node_weight: np.ndarray
    node_word: np.ndarray
    k: int
    L: int
    weighting: int = 0
    scoring: int = 0

    @property
    def n_nodes(self) -> int:
        return len(self.node_word)

    @property
    def n_words(self) -> int:
        return int(self.node_word.max()) + 1 if len(self.node_word) else 0


def make_vocabulary(seed: int, k: int = 10, L: int = 3, weighting: int = 0, flip_p: float = 0.12, ragged: float = 0.1,
                    stop_frac: float = 0.02) -> FlatVoc:
    """Random hierarchical vocabulary shaped like DBoW2's k-means tree: children are noisy copies of their
    parent, ids are assigned level by level within a parent (children contiguous, larger than the parent, as
    HKmeansStep creates them); `ragged` of the inner nodes get fewer than k children or stay leaves above depth
    L; leaves get idf-like weights log(N / n_i), `stop_frac` of them weight 0 (stopped words)."""
    rng = np.random.default_rng(seed)
    desc = [rand_desc(rng, 1)[0]]
    children = [[]]
    depth = [0]

    def grow(node):
        if depth[node] >= L:
            return
        if node != 0 and rng.random() < ragged / 2:
            return                                   # a leaf above the last level
        nk = k if rng.random() >= ragged else int(rng.integers(1, k + 1))
        first = len(desc)
        for _ in range(nk):
            desc.append(flip_bits(rng, desc[node][None, :], flip_p)[0])
            children.append([])
            depth.append(depth[node] + 1)
        children[node] = list(range(first, first + nk))
        for c in children[node]:
            grow(c)

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    grow(0)
    n = len(desc)
    child_start = np.zeros(n + 1, np.int32)
    child_start[1:] = np.cumsum([len(c) for c in children])
    child_ids = np.array([c for cs in children for c in cs], np.int32)
    node_word = np.full(n, -1, np.int32)
    leaves = [i for i in range(1, n) if not children[i]]
    node_word[leaves] = np.arange(len(leaves), dtype=np.int32)
    node_weight = np.zeros(n, np.float64)
    w = np.log(1000.0 / rng.integers(1, 400, len(leaves)))
    if weighting in (1, 3):                          # TF / BINARY: weight 1 (setNodeWeights, :984-997)
        w[:] = 1.0
    w[rng.random(len(leaves)) < stop_frac] = 0.0
    node_weight[leaves] = w
    return FlatVoc(child_start, child_ids, np.ascontiguousarray(np.stack(desc), np.uint8), node_weight, node_word, k, L,
                   weighting, 0)


def vocabulary_features(seed: int, voc: FlatVoc, n: int, flip_p: float = 0.1) -> np.ndarray:
    """n descriptors that are noisy copies of random leaves of the vocabulary (so that images share words)."""
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(voc.node_word >= 0)[0]
    if n == 0 or len(leaves) == 0:
        return np.zeros((0, 32), np.uint8)
    return flip_bits(rng, voc.node_desc[rng.choice(leaves, n)], flip_p)


# ---- config 3: offline replay of many stereo frames (vectorised over frames) ---------------------

@dataclass
class Replay:
    """Arenas + job tables of an n_frames replay: per frame one stereo matchGrid(points), one stereo
    matchGrid(lines), one temporal match(points) and one temporal match(lines) against the previous
    frame (app/plslam_dataset.cpp:114-172 -> stereoFrame.cpp:157,356 and stereoFrameHandler.cpp:168,191).
    Descriptor arena row layout per frame: [left points | right points | left lines | right lines]."""
    n_frames: int
    arena: np.ndarray            # rows x 32 uint8
    n_pts: np.ndarray            # per frame
    n_lines: np.ndarray
    off_pl: np.ndarray           # arena row offsets per frame
    off_pr: np.ndarray
    off_ll: np.ndarray
    off_lr: np.ndarray
    coords: np.ndarray           # int32 arena: per frame [xy of left points | xyxy of left lines]
    off_cpts: np.ndarray
    off_clines: np.ndarray
    cell_start: np.ndarray       # int32 arena: per frame [points grid | lines grid], 3073 each
    cell_items: np.ndarray
    off_items_p: np.ndarray
    off_items_l: np.ndarray
    dirs2: np.ndarray            # float64 arena: per frame 2 * n_lines
    off_dirs: np.ndarray
    kp_l: np.ndarray             # float32 pixel coordinates (for the stereo gates), same order as arena
    kp_r: np.ndarray
    ln_l: np.ndarray
    ln_r: np.ndarray
    off_m_p: np.ndarray          # match-vector arena offsets
    off_m_l: np.ndarray
    n_m: int


def make_replay(seed: int, n_frames: int, mean_pts: int = 600, sd_pts: int = 50, mean_lines: int = 200,
                sd_lines: int = 25, flip_p: float = 0.08, outlier_frac: float = 0.25) -> Replay:
    rng = np.random.default_rng(seed)
    F = n_frames
    n_pts = np.clip(np.rint(rng.normal(mean_pts, sd_pts, F)), 2, None).astype(np.int64)
    n_lines = np.clip(np.rint(rng.normal(mean_lines, sd_lines, F)), 2, None).astype(np.int64)
    per_frame = 2 * n_pts + 2 * n_lines
    base = np.concatenate([[0], np.cumsum(per_frame)])
    off_pl = base[:-1]
    off_pr = off_pl + n_pts
    off_ll = off_pr + n_pts
    off_lr = off_ll + n_lines
    arena = np.empty((int(base[-1]), 32), np.uint8)
    P, Lm = int(n_pts.max()), int(n_lines.max())
    b = 19.0

    def evolve(prev, n, cap):
        """Left descriptors of the next frame: noisy copies of the previous frame's, shuffled, with
        outlier_frac replaced by fresh random rows (so the temporal matcher has true matches)."""
        cur = rand_desc(rng, cap)
        k = min(len(prev), n)
        cur[:k] = flip_bits(rng, prev[:k], flip_p / 2)
        out = rng.random(cap) < outlier_frac
        cur[out] = rand_desc(rng, int(out.sum()))
        return cur[rng.permutation(cap)][:n]

    kp_l = np.empty((int(n_pts.sum()), 2), np.float32)
    kp_r = np.empty_like(kp_l)
    ln_l = np.empty((int(n_lines.sum()), 4), np.float32)
    ln_r = np.empty_like(ln_l)
    pbase = np.concatenate([[0], np.cumsum(n_pts)])
    lbase = np.concatenate([[0], np.cumsum(n_lines)])
    prev_p = rand_desc(rng, P)
    prev_l = rand_desc(rng, Lm)
    for f in range(F):
        n, m = int(n_pts[f]), int(n_lines[f])
        dl = evolve(prev_p, n, max(n, len(prev_p)))
        prev_p = dl
        dr = flip_bits(rng, dl, flip_p)
        out = rng.random(n) < outlier_frac
        dr[out] = rand_desc(rng, int(out.sum()))
        perm = rng.permutation(n)
        arena[off_pl[f]:off_pl[f] + n] = dl
        arena[off_pr[f]:off_pr[f] + n] = dr[perm]
        xl = rng.uniform(b, IMG_W - b, n); yl = rng.uniform(b, IMG_H - b, n)
        xr = xl - rng.uniform(1.0, 100.0, n); yr = yl + rng.normal(0.0, 0.5, n)
        xr[out] = rng.uniform(b, IMG_W - b, int(out.sum())); yr[out] = rng.uniform(b, IMG_H - b, int(out.sum()))
        kp_l[pbase[f]:pbase[f + 1]] = np.stack([xl, yl], 1)
        kp_r[pbase[f]:pbase[f + 1]] = np.stack([xr, yr], 1)[perm]

        ll = evolve(prev_l, m, max(m, len(prev_l)))
        prev_l = ll
        lr = flip_bits(rng, ll, flip_p)
        lout = rng.random(m) < outlier_frac
        lr[lout] = rand_desc(rng, int(lout.sum()))
        lperm = rng.permutation(m)
        arena[off_ll[f]:off_ll[f] + m] = ll
        arena[off_lr[f]:off_lr[f] + m] = lr[lperm]
        sx = rng.uniform(b, IMG_W - b, m); sy = rng.uniform(b, IMG_H - b, m)
        ang = rng.uniform(0, 2 * np.pi, m); length = rng.uniform(0.025 * IMG_H, 150.0, m)
        ex = np.clip(sx + length * np.cos(ang), 1.0, IMG_W - 2.0); ey = np.clip(sy + length * np.sin(ang), 1.0, IMG_H - 2.0)
        d = rng.uniform(1.0, 100.0, m)
        r = np.stack([sx - d + rng.normal(0, 0.3, m), sy + rng.normal(0, 0.5, m),
                      ex - d + rng.normal(0, 0.3, m), ey + rng.normal(0, 0.5, m)], 1)
        no = int(lout.sum())
        r[lout] = np.stack([rng.uniform(b, IMG_W - b, no), rng.uniform(b, IMG_H - b, no),
                            rng.uniform(b, IMG_W - b, no), rng.uniform(b, IMG_H - b, no)], 1)
        ln_l[lbase[f]:lbase[f + 1]] = np.stack([sx, sy, ex, ey], 1)
        ln_r[lbase[f]:lbase[f + 1]] = r[lperm]

    # query coordinates (stereoFrame.cpp:140-143 / :329-333): trunc(px * inv)
    kl = kp_l.astype(np.float64)
    xy = np.stack([np.trunc(kl[:, 0] * INV_W), np.trunc(kl[:, 1] * INV_H)], 1).astype(np.int32)
    l64 = ln_l.astype(np.float64)
    xyxy = np.stack([np.trunc(l64[:, 0] * INV_W), np.trunc(l64[:, 1] * INV_H),
                     np.trunc(l64[:, 2] * INV_W), np.trunc(l64[:, 3] * INV_H)], 1).astype(np.int32)
    ccount = 2 * n_pts + 4 * n_lines
    cbase = np.concatenate([[0], np.cumsum(ccount)])
    off_cpts = cbase[:-1]
    off_clines = off_cpts + 2 * n_pts
    coords = np.empty(int(cbase[-1]), np.int32)
    for f in range(F):
        coords[off_cpts[f]:off_cpts[f] + 2 * n_pts[f]] = xy[pbase[f]:pbase[f + 1]].ravel()
        coords[off_clines[f]:off_clines[f] + 4 * n_lines[f]] = xyxy[lbase[f]:lbase[f + 1]].ravel()

    # grids over the RIGHT features, all frames at once: key = frame * n_cells + x * rows + y
    n_cells = G.GRID_ROWS * G.GRID_COLS
    kr = kp_r.astype(np.float64)
    cx = np.trunc(kr[:, 0] * INV_W).astype(np.int64); cy = np.trunc(kr[:, 1] * INV_H).astype(np.int64)
    fid = np.repeat(np.arange(F), n_pts)
    local = np.arange(len(kr)) - np.repeat(pbase[:-1], n_pts)
    ok = (cx >= 0) & (cx < G.GRID_COLS) & (cy >= 0) & (cy < G.GRID_ROWS)
    key_p = fid[ok] * n_cells + cx[ok] * G.GRID_ROWS + cy[ok]
    order = np.argsort(key_p, kind="stable")
    items_p = local[ok][order].astype(np.int32)
    cnt_p = np.bincount(key_p, minlength=F * n_cells).reshape(F, n_cells)

    r64 = ln_r.astype(np.float64)
    ids, lcx, lcy = G.line_cells(r64[:, 0] * INV_W, r64[:, 1] * INV_H, r64[:, 2] * INV_W, r64[:, 3] * INV_H)
    lfid = np.repeat(np.arange(F), n_lines)[ids]
    llocal = ids - np.repeat(lbase[:-1], n_lines)[ids]
    okl = (lcx >= 0) & (lcx < G.GRID_COLS) & (lcy >= 0) & (lcy < G.GRID_ROWS)
    key_l = lfid[okl] * n_cells + lcx[okl] * G.GRID_ROWS + lcy[okl]
    order_l = np.argsort(key_l, kind="stable")
    items_l = llocal[okl][order_l].astype(np.int32)
    cnt_l = np.bincount(key_l, minlength=F * n_cells).reshape(F, n_cells)

    cell_start = np.zeros((F, 2, n_cells + 1), np.int32)
    np.cumsum(cnt_p, axis=1, out=cell_start[:, 0, 1:])
    np.cumsum(cnt_l, axis=1, out=cell_start[:, 1, 1:])
    tot_p = cell_start[:, 0, -1].astype(np.int64); tot_l = cell_start[:, 1, -1].astype(np.int64)
    ibase = np.concatenate([[0], np.cumsum(tot_p + tot_l)])
    off_items_p = ibase[:-1]
    off_items_l = off_items_p + tot_p
    cell_items = np.empty(int(ibase[-1]), np.int32)
    pstart = np.concatenate([[0], np.cumsum(tot_p)]); lstart = np.concatenate([[0], np.cumsum(tot_l)])
    for f in range(F):
        cell_items[off_items_p[f]:off_items_p[f] + tot_p[f]] = items_p[pstart[f]:pstart[f + 1]]
        cell_items[off_items_l[f]:off_items_l[f] + tot_l[f]] = items_l[lstart[f]:lstart[f + 1]]

    fr = ln_r
    vx = (fr[:, 2] - fr[:, 0]).astype(np.float64) * INV_W
    vy = (fr[:, 3] - fr[:, 1]).astype(np.float64) * INV_H
    with np.errstate(invalid="ignore", divide="ignore"):
        mag = np.sqrt(vx * vx + vy * vy)
        dirs2 = np.stack([vx / mag, vy / mag], 1).ravel()
    off_dirs = 2 * lbase[:-1]

    mbase = np.concatenate([[0], np.cumsum(n_pts + n_lines)])
    off_m_p = mbase[:-1]
    off_m_l = off_m_p + n_pts
    return Replay(F, arena, n_pts, n_lines, off_pl, off_pr, off_ll, off_lr, coords, off_cpts, off_clines,
                  cell_start.reshape(-1), cell_items, off_items_p, off_items_l, dirs2, off_dirs, kp_l, kp_r, ln_l, ln_r,
                  off_m_p, off_m_l, int(mbase[-1]))
