"""ctypes front-ends for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``oracle.port``  -- oracle/plm_oracle.c, the plain-C restatement (``_ref/libploracle.so``)
* ``oracle.ref``   -- the reference itself, /root/reference/stvo-pl/src/{matching,gridStructure,
  lineIterator}.cpp and /root/reference/src/mapFeatures.cpp compiled unmodified
  (``_ref/libplref.so``; see oracle/Makefile)
* ``oracle.ref_stereo`` -- the reference's stereo drivers and gates, stvo-pl/src/{stereoFrame,stereoFeatures,
  pinholeStereoCamera}.cpp compiled unmodified (``_ref/libplref_stereo.so``)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU arms may import this package; the
product (pl_inertial_slam_b200) never does.  Nothing here reads /root/reference at run time: the
GPU box only has the prebuilt ``_ref/*.so`` files.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference itself."""
    targets = ["oracle"]
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "stvo-pl", "src")):
        targets.append("ref")
        if os.path.exists(os.path.join(_HERE, "..", "pl_inertial_slam_b200", "lib", "libplmatch.so")):
            targets += ["stvo_gpu", "map_gpu", "dbow_gpu"]  # the C++ drop-ins over the CUDA library, same harnesses
    out = subprocess.run(["make", "-C", _HERE] + targets, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _load(name: str):
    path = os.path.join(_REF_DIR, name)
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return C.CDLL(path)


def _desc(a: np.ndarray):
    """(pointer, rows, step) for an n x 32 uint8 matrix whose rows are contiguous."""
    assert a.dtype == np.uint8 and a.ndim == 2 and a.shape[1] == 32, (a.dtype, a.shape)
    assert a.strides[1] == 1
    return a.ctypes.data_as(_u8p), int(a.shape[0]), C.c_size_t(a.strides[0] if a.shape[0] > 1 else 32)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


class _Port:
    """oracle/plm_oracle.c"""

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            L = _load("libploracle.so")
            L.plo_hamming256.restype = C.c_int
            L.plo_hamming256.argtypes = [_u8p, _u8p]
            L.plo_knn2.restype = None
            L.plo_knn2.argtypes = [_u8p, C.c_int, C.c_size_t, _u8p, C.c_int, C.c_size_t, _i32p, _i32p]
            L.plo_match_nnr.restype = C.c_int
            L.plo_match_nnr.argtypes = [_u8p, C.c_int, C.c_size_t, _u8p, C.c_int, C.c_size_t, C.c_float, _i32p]
            L.plo_match.restype = C.c_int
            L.plo_match.argtypes = [_u8p, C.c_int, C.c_size_t, _u8p, C.c_int, C.c_size_t, C.c_float, C.c_int, _i32p]
            L.plo_line_coords.restype = C.c_int
            L.plo_line_coords.argtypes = [C.c_double] * 4 + [_i32p, C.c_int]
            L.plo_match_grid_points.restype = C.c_int
            L.plo_match_grid_points.argtypes = [_i32p, _u8p, C.c_int, C.c_size_t, _i32p, _i32p, C.c_int, C.c_int,
                                                _u8p, C.c_int, C.c_size_t, _i32p, C.c_double, C.c_int, _i32p]
            L.plo_match_grid_lines.restype = C.c_int
            L.plo_match_grid_lines.argtypes = [_i32p, _u8p, C.c_int, C.c_size_t, _i32p, _i32p, C.c_int, C.c_int,
                                               _u8p, C.c_int, C.c_size_t, _f64p, C.c_double, _i32p, C.c_double,
                                               C.c_int, _i32p]
            L.plo_match_grid_shard.restype = C.c_int
            L.plo_match_grid_shard.argtypes = [C.c_int, _i32p, _u8p, C.c_int, C.c_size_t, C.c_int64, _i32p, _i32p,
                                               C.c_int, C.c_int, _u8p, C.c_int, C.c_size_t, _f64p, C.c_double, _i32p,
                                               C.c_double, C.c_int, _i32p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.plo_stereo_filter_points.restype = C.c_int
            L.plo_stereo_filter_points.argtypes = [_f32p, _f32p, _i32p, C.c_int, C.c_double, C.c_double, _u8p, _f64p]
            L.plo_stereo_filter_lines.restype = C.c_int
            L.plo_stereo_filter_lines.argtypes = [_f32p, _f32p, _i32p, C.c_int, C.c_double, C.c_double, C.c_double,
                                                  C.c_double, _u8p, _f64p]
            L.plo_line_overlap_stereo.restype = C.c_double
            L.plo_line_overlap_stereo.argtypes = [C.c_double] * 5
            L.plo_csr_from_points.restype = C.c_int
            L.plo_csr_from_points.argtypes = [_f32p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _i32p, _i32p]
            L.plo_csr_from_lines.restype = C.c_int
            L.plo_csr_from_lines.argtypes = [_f32p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _i32p, _i32p, _f64p]
            L.plo_stereo_points.restype = C.c_int
            L.plo_stereo_points.argtypes = [_f32p, _u8p, C.c_int, _f32p, _u8p, C.c_int, C.c_double, C.c_double, C.c_int,
                                            C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _f64p, _i32p,
                                            _i32p, _f64p, _f64p]
            L.plo_stereo_lines.restype = C.c_int
            L.plo_stereo_lines.argtypes = [_f32p, _u8p, C.c_int, _f32p, _u8p, C.c_int, C.c_double, C.c_double, C.c_int,
                                           C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                           C.c_double, C.c_double, _f64p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]
            u32p = C.POINTER(C.c_uint32)
            L.plo_bow_word.restype = C.c_int
            L.plo_bow_word.argtypes = [_i32p, _i32p, _u8p, _u8p]
            L.plo_bow_transform.restype = C.c_int
            L.plo_bow_transform.argtypes = [_i32p, _i32p, _u8p, _f64p, _i32p, C.c_int, _u8p, C.c_int, C.c_size_t, u32p, _f64p]
            L.plo_bow_score.restype = C.c_double
            L.plo_bow_score.argtypes = [u32p, _f64p, C.c_int, u32p, _f64p, C.c_int]
            L.plo_line_pair_filter.restype = C.c_int
            L.plo_line_pair_filter.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_double, C.c_double, _u8p, _f64p, _f64p]
            L.plo_med_desc.restype = None
            L.plo_med_desc.argtypes = [_u8p, C.c_size_t, _f64p, _i32p, C.c_int, _i32p, _u8p, _f64p]
            self._lib = L
        return self._lib

    def bow_transform(self, fv, desc):
        """TemplatedVocabulary::transform(features, BowVector) on a FlatVocabulary -> (ids uint32, vals float64)."""
        u32p = C.POINTER(C.c_uint32)
        p, n, s = _desc(np.ascontiguousarray(desc, np.uint8).reshape(-1, 32))
        ids, vals = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
        m = self.lib.plo_bow_transform(fv.child_start.ctypes.data_as(_i32p), fv.child_ids.ctypes.data_as(_i32p),
                                       fv.node_desc.ctypes.data_as(_u8p), fv.node_weight.ctypes.data_as(_f64p),
                                       fv.node_word.ctypes.data_as(_i32p), fv.weighting, p, n, s,
                                       ids.ctypes.data_as(u32p), vals.ctypes.data_as(_f64p))
        return ids[:m].copy(), vals[:m].copy()

    def bow_score(self, v1, v2) -> float:
        """L1Scoring::score(v1, v2)."""
        u32p = C.POINTER(C.c_uint32)
        i1, x1 = np.ascontiguousarray(v1[0], np.uint32), np.ascontiguousarray(v1[1], np.float64)
        i2, x2 = np.ascontiguousarray(v2[0], np.uint32), np.ascontiguousarray(v2[1], np.float64)
        return self.lib.plo_bow_score(i1.ctypes.data_as(u32p), x1.ctypes.data_as(_f64p), len(i1),
                                      i2.ctypes.data_as(u32p), x2.ctypes.data_as(_f64p), len(i2))

    def line_pair_filter(self, ln1, ln2, m12, overlap_th=0.75, line_sim_th=0.75):
        """lineSegmentOverlap + direction similarity per matched line pair -> (n_kept, keep, overlap, sim)."""
        ln1 = np.ascontiguousarray(ln1, np.float32).reshape(-1, 4)
        ln2 = np.ascontiguousarray(ln2, np.float32).reshape(-1, 4)
        m, mp = _i32(m12)
        n1 = len(m)
        keep = np.zeros(max(n1, 1), np.uint8)
        ov, sim = np.zeros(max(n1, 1), np.float64), np.zeros(max(n1, 1), np.float64)
        n = self.lib.plo_line_pair_filter(ln1.ctypes.data_as(_f32p), n1, ln2.ctypes.data_as(_f32p), len(ln2), mp,
                                          overlap_th, line_sim_th, keep.ctypes.data_as(_u8p), ov.ctypes.data_as(_f64p),
                                          sim.ctypes.data_as(_f64p))
        return n, keep[:n1], ov[:n1], sim[:n1]

    # ---- local-map selection / reprojection gates (mapHandler.cpp:583-682, :685-803) -------------------------
    @staticmethod
    def _view(T, cam, inv_w, inv_h, width, height):
        class V(C.Structure):
            _fields_ = [("T", C.c_double * 12), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                        ("inv_width", C.c_double), ("inv_height", C.c_double), ("width", C.c_int32), ("height", C.c_int32)]
        v = V()
        v.T[:] = [float(x) for x in np.asarray(T, np.float64).reshape(-1)[:12]]
        v.fx, v.fy, v.cx, v.cy = [float(x) for x in cam]
        v.inv_width, v.inv_height, v.width, v.height = float(inv_w), float(inv_h), int(width), int(height)
        return v

    def map_select(self, X, active, T, cam, inv_w, inv_h, width, height):
        """-> (sel, coords, pf) of the landmarks projected inside the image, in order."""
        X = np.ascontiguousarray(X, np.float64)
        is_lines = X.shape[1] == 6
        n, per = len(X), 2 if is_lines else 1
        act = None if active is None else np.ascontiguousarray(active, np.uint8)
        sel, coords, pf = np.zeros(max(n, 1), np.int32), np.zeros((max(n, 1), 2 * per), np.int32), np.zeros((max(n, 1), 2 * per))
        v = self._view(T, cam, inv_w, inv_h, width, height)
        self.lib.plo_map_select.restype = C.c_int
        m = self.lib.plo_map_select(int(is_lines), X.ctypes.data_as(_f64p), None if act is None else act.ctypes.data_as(_u8p), n,
                                    C.byref(v), sel.ctypes.data_as(_i32p), coords.ctypes.data_as(_i32p), pf.ctypes.data_as(_f64p))
        return sel[:m].copy(), coords[:m].copy(), pf[:m].copy()

    def map_gate(self, pf, m12, feat, max_epip, count):
        pf = np.ascontiguousarray(pf, np.float64)
        is_lines = pf.shape[1] == 4
        m, mp = _i32(m12)
        feat = np.ascontiguousarray(feat, np.float64)
        ok = np.zeros(max(len(m), 1), np.uint8)
        self.lib.plo_map_gate.restype = C.c_int
        self.lib.plo_map_gate.argtypes = [C.c_int, _f64p, _i32p, C.c_int, _f64p, C.c_int, C.c_double, _u8p, C.c_int]
        c = self.lib.plo_map_gate(int(is_lines), pf.ctypes.data_as(_f64p), mp, len(m), feat.ctypes.data_as(_f64p), len(feat),
                                  float(max_epip), ok.ctypes.data_as(_u8p), int(count))
        return c, ok[:len(m)].copy()

    def med_desc(self, desc, dirs, obs_start):
        """MapPoint / MapLine::updateAverageDescDir over a batch of landmarks (src/mapFeatures.cpp:51-93,
        :121-163) -> (med_idx int32[n_lm], med_desc uint8[n_lm, 32], med_dir float64[n_lm, 3])."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        obs_start, osp = _i32(obs_start)
        n_lm = len(obs_start) - 1
        dirs_c = None if dirs is None else np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        med_idx = np.empty(n_lm, np.int32)
        med = np.zeros((n_lm, 32), np.uint8)
        med_dir = np.zeros((n_lm, 3), np.float64)
        self.lib.plo_med_desc(desc.ctypes.data_as(_u8p), 32, None if dirs_c is None else dirs_c.ctypes.data_as(_f64p),
                              osp, n_lm, med_idx.ctypes.data_as(_i32p), med.ctypes.data_as(_u8p),
                              med_dir.ctypes.data_as(_f64p))
        return med_idx, med, med_dir

    def distance(self, a, b) -> int:
        a = np.ascontiguousarray(a, np.uint8).reshape(32)
        b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return self.lib.plo_hamming256(a.ctypes.data_as(_u8p), b.ctypes.data_as(_u8p))

    def knn2(self, d1, d2):
        """-> (idx n1x2 int32, dist n1x2 int32); absent slots are (-1, INT_MAX)."""
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        idx = np.empty((n1, 2), np.int32)
        dist = np.empty((n1, 2), np.int32)
        self.lib.plo_knn2(p1, n1, s1, p2, n2, s2, idx.ctypes.data_as(_i32p), dist.ctypes.data_as(_i32p))
        return idx, dist

    def knn2_packed(self, d1, d2, idx_base: int = 0) -> np.ndarray:
        """Packed (dist << 32 | idx) uint64 keys, n1 x 2, UINT64_MAX where absent."""
        idx, dist = self.knn2(d1, d2)
        key = (dist.astype(np.uint64) << np.uint64(32)) | (idx.astype(np.int64) + idx_base).astype(np.uint64)
        key[idx < 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
        return key

    def match_nnr(self, d1, d2, nnr, m12=None):
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = self.lib.plo_match_nnr(p1, n1, s1, p2, n2, s2, C.c_float(nnr), m12.ctypes.data_as(_i32p))
        return n, m12

    def match(self, d1, d2, nnr, best_lr=True, m12=None):
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = self.lib.plo_match(p1, n1, s1, p2, n2, s2, C.c_float(nnr), int(bool(best_lr)),
                               m12.ctypes.data_as(_i32p))
        return n, m12

    def line_coords(self, x1, y1, x2, y2, max_cells=4096) -> np.ndarray:
        cells = np.empty((max_cells, 2), np.int32)
        n = self.lib.plo_line_coords(x1, y1, x2, y2, cells.ctypes.data_as(_i32p), max_cells)
        assert n <= max_cells
        return cells[:n].copy()

    def match_grid_points(self, xy, d1, cell_start, cell_items, rows, cols, d2, win, ratio, best_lr=True, m12=None):
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        xy, xyp = _i32(xy)
        cs, csp = _i32(cell_start)
        ci, cip = _i32(cell_items if len(cell_items) else np.zeros(1, np.int32))
        w, wp = _i32(win)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = self.lib.plo_match_grid_points(xyp, p1, n1, s1, csp, cip, rows, cols, p2, n2, s2, wp,
                                           float(ratio), int(bool(best_lr)), m12.ctypes.data_as(_i32p))
        return n, m12

    def match_grid_lines(self, xyxy, d1, cell_start, cell_items, rows, cols, d2, dirs2, line_sim_th, win, ratio,
                         best_lr=True, m12=None):
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        xy, xyp = _i32(xyxy)
        cs, csp = _i32(cell_start)
        ci, cip = _i32(cell_items if len(cell_items) else np.zeros(1, np.int32))
        w, wp = _i32(win)
        dirs2 = np.ascontiguousarray(dirs2, np.float64)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = self.lib.plo_match_grid_lines(xyp, p1, n1, s1, csp, cip, rows, cols, p2, n2, s2,
                                          dirs2.ctypes.data_as(_f64p), float(line_sim_th), wp, float(ratio),
                                          int(bool(best_lr)), m12.ctypes.data_as(_i32p))
        return n, m12

    def match_grid_shard(self, is_lines, coords, d1, i1_base, cell_start, cell_items, rows, cols, d2, dirs2,
                         line_sim_th, win, ratio, best_lr, m12, seed=None):
        """Row-shard form: (accepts, m12, colmin uint16[n2], m21key uint64[n2]); no mutual check."""
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        xy, xyp = _i32(coords)
        cs, csp = _i32(cell_start)
        ci, cip = _i32(cell_items if len(cell_items) else np.zeros(1, np.int32))
        w, wp = _i32(win)
        dirs = np.ascontiguousarray(dirs2 if dirs2 is not None else np.zeros((max(n2, 1), 2)), np.float64)
        m12 = np.array(m12, np.int32)
        colmin = np.full(max(n2, 1), 0xFFFF, np.uint16)
        key = np.full(max(n2, 1), 0xFFFFFFFFFFFFFFFF, np.uint64)
        seed_p = None
        if seed is not None:
            seed = np.ascontiguousarray(seed, np.uint16)
            seed_p = seed.ctypes.data_as(C.c_void_p)
        n = self.lib.plo_match_grid_shard(int(bool(is_lines)), xyp, p1, n1, s1, int(i1_base), csp, cip, rows, cols, p2,
                                          n2, s2, dirs.ctypes.data_as(_f64p), float(line_sim_th), wp, float(ratio),
                                          int(bool(best_lr)), m12.ctypes.data_as(_i32p), seed_p,
                                          colmin.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p))
        return n, m12, colmin[:n2], key[:n2]

    def stereo_filter_points(self, kp_l, kp_r, m12, max_dist_epip=1.0, min_disp=1.0):
        kp_l = np.ascontiguousarray(kp_l, np.float32)
        kp_r = np.ascontiguousarray(kp_r, np.float32)
        m, mp = _i32(m12)
        n1 = len(m)
        keep = np.zeros(n1, np.uint8)
        disp = np.zeros(n1, np.float64)
        n = self.lib.plo_stereo_filter_points(kp_l.ctypes.data_as(_f32p), kp_r.ctypes.data_as(_f32p), mp, n1,
                                              max_dist_epip, min_disp, keep.ctypes.data_as(_u8p),
                                              disp.ctypes.data_as(_f64p))
        return n, keep, disp

    def stereo_filter_lines(self, ln_l, ln_r, m12, min_disp=1.0, line_horiz_th=0.1, stereo_overlap_th=0.75,
                            ls_min_disp_ratio=0.7):
        ln_l = np.ascontiguousarray(ln_l, np.float32)
        ln_r = np.ascontiguousarray(ln_r, np.float32)
        m, mp = _i32(m12)
        n1 = len(m)
        keep = np.zeros(n1, np.uint8)
        disp = np.zeros((n1, 2), np.float64)
        n = self.lib.plo_stereo_filter_lines(ln_l.ctypes.data_as(_f32p), ln_r.ctypes.data_as(_f32p), mp, n1,
                                             min_disp, line_horiz_th, stereo_overlap_th, ls_min_disp_ratio,
                                             keep.ctypes.data_as(_u8p), disp.ctypes.data_as(_f64p))
        return n, keep, disp


    # ---- whole stereo drivers (stereoFrame.cpp:131-184, :320-409) ---------------------------------------
    def csr_from_points(self, kp, inv_w, inv_h, rows=48, cols=64):
        kp = np.ascontiguousarray(kp, np.float32).reshape(-1, 2)
        cs = np.zeros(rows * cols + 1, np.int32)
        ci = np.zeros(max(1, len(kp)), np.int32)
        n = self.lib.plo_csr_from_points(kp.ctypes.data_as(_f32p), len(kp), inv_w, inv_h, rows, cols,
                                         cs.ctypes.data_as(_i32p), ci.ctypes.data_as(_i32p))
        return cs, ci[:n]

    def csr_from_lines(self, ln, inv_w, inv_h, rows=48, cols=64):
        ln = np.ascontiguousarray(ln, np.float32).reshape(-1, 4)
        cs = np.zeros(rows * cols + 1, np.int32)
        dirs = np.zeros((len(ln), 2), np.float64)
        n = self.lib.plo_csr_from_lines(ln.ctypes.data_as(_f32p), len(ln), inv_w, inv_h, rows, cols,
                                        cs.ctypes.data_as(_i32p), None, None)
        ci = np.zeros(max(1, n), np.int32)
        self.lib.plo_csr_from_lines(ln.ctypes.data_as(_f32p), len(ln), inv_w, inv_h, rows, cols,
                                    cs.ctypes.data_as(_i32p), ci.ctypes.data_as(_i32p), dirs.ctypes.data_as(_f64p))
        return cs, ci[:n], dirs

    def stereo_points(self, kp_l, d_l, kp_r, d_r, inv_w, inv_h, cam, rows=48, cols=64, matching_s_ws=10, ratio=0.9,
                      best_lr=True, max_dist_epip=1.0, min_disp=1.0):
        """StereoFrame::matchStereoPoints -> dict(m12, kept_i1, disp, P)."""
        kp_l = np.ascontiguousarray(kp_l, np.float32).reshape(-1, 2)
        kp_r = np.ascontiguousarray(kp_r, np.float32).reshape(-1, 2)
        d_l = np.ascontiguousarray(d_l, np.uint8).reshape(-1, 32)
        d_r = np.ascontiguousarray(d_r, np.uint8).reshape(-1, 32)
        n_l, n_r = len(kp_l), len(kp_r)
        cam = np.ascontiguousarray(cam, np.float64)
        m12 = np.full(n_l, -1, np.int32)
        kept = np.zeros(max(1, n_l), np.int32)
        disp = np.zeros(max(1, n_l), np.float64)
        P = np.zeros((max(1, n_l), 3), np.float64)
        n = self.lib.plo_stereo_points(kp_l.ctypes.data_as(_f32p), d_l.ctypes.data_as(_u8p), n_l,
                                       kp_r.ctypes.data_as(_f32p), d_r.ctypes.data_as(_u8p), n_r, inv_w, inv_h, rows,
                                       cols, matching_s_ws, ratio, int(best_lr), max_dist_epip, min_disp,
                                       cam.ctypes.data_as(_f64p), m12.ctypes.data_as(_i32p), kept.ctypes.data_as(_i32p),
                                       disp.ctypes.data_as(_f64p), P.ctypes.data_as(_f64p))
        return dict(m12=m12, kept_i1=kept[:n].copy(), disp=disp[:n].copy(), P=P[:n].copy())

    def stereo_lines(self, ln_l, d_l, ln_r, d_r, inv_w, inv_h, cam, rows=48, cols=64, matching_s_ws=10, ratio=0.9,
                     line_sim_th=0.75, best_lr=True, min_disp=1.0, line_horiz_th=0.1, stereo_overlap_th=0.75,
                     ls_min_disp_ratio=0.7):
        """StereoFrame::matchStereoLines -> dict(m12, kept_i1, disp_se, sP, eP, le)."""
        ln_l = np.ascontiguousarray(ln_l, np.float32).reshape(-1, 4)
        ln_r = np.ascontiguousarray(ln_r, np.float32).reshape(-1, 4)
        d_l = np.ascontiguousarray(d_l, np.uint8).reshape(-1, 32)
        d_r = np.ascontiguousarray(d_r, np.uint8).reshape(-1, 32)
        n_l, n_r = len(ln_l), len(ln_r)
        cam = np.ascontiguousarray(cam, np.float64)
        m12 = np.full(n_l, -1, np.int32)
        cap = max(1, n_l)
        kept = np.zeros(cap, np.int32)
        dse = np.zeros((cap, 2), np.float64)
        sP, eP, le = (np.zeros((cap, 3), np.float64) for _ in range(3))
        n = self.lib.plo_stereo_lines(ln_l.ctypes.data_as(_f32p), d_l.ctypes.data_as(_u8p), n_l,
                                      ln_r.ctypes.data_as(_f32p), d_r.ctypes.data_as(_u8p), n_r, inv_w, inv_h, rows,
                                      cols, matching_s_ws, ratio, line_sim_th, int(best_lr), min_disp, line_horiz_th,
                                      stereo_overlap_th, ls_min_disp_ratio, cam.ctypes.data_as(_f64p),
                                      m12.ctypes.data_as(_i32p), kept.ctypes.data_as(_i32p), dse.ctypes.data_as(_f64p),
                                      sP.ctypes.data_as(_f64p), eP.ctypes.data_as(_f64p), le.ctypes.data_as(_f64p))
        return dict(m12=m12, kept_i1=kept[:n].copy(), disp_se=dse[:n].copy(), sP=sP[:n].copy(), eP=eP[:n].copy(),
                    le=le[:n].copy())


class _Ref:
    """The reference's own matching.cpp / gridStructure.cpp / lineIterator.cpp (compiled unmodified),
    or -- with libname="libstvo_gpu.so" -- the product's C++ drop-in for matching.cpp (StVO::
    signatures over the CUDA library) behind the same extern "C" harness."""

    def __init__(self, libname: str = "libplref.so"):
        self._lib = None
        self._libname = libname

    def available(self) -> bool:
        try:
            return self.lib is not None
        except (FileNotFoundError, OSError, RuntimeError):
            return False

    @property
    def lib(self):
        if self._lib is None:
            L = _load(self._libname)
            L.plref_set_threads.argtypes = [C.c_int]
            L.plref_get_threads.restype = C.c_int
            L.plref_set_config.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
            L.plref_distance.restype = C.c_int
            L.plref_distance.argtypes = [_u8p, _u8p]
            sig = [_u8p, C.c_int, C.c_size_t, _u8p, C.c_int, C.c_size_t, C.c_float, _i32p, C.POINTER(C.c_int)]
            L.plref_match_nnr.restype = C.c_int
            L.plref_match_nnr.argtypes = sig
            L.plref_match.restype = C.c_int
            L.plref_match.argtypes = sig
            L.plref_match_grid_points.restype = C.c_int
            L.plref_match_grid_points.argtypes = [_i32p, _u8p, C.c_int, C.c_size_t, _i32p, _i32p, C.c_int, C.c_int,
                                                  _u8p, C.c_int, C.c_size_t, _i32p, _i32p, C.POINTER(C.c_int)]
            L.plref_match_grid_lines.restype = C.c_int
            L.plref_match_grid_lines.argtypes = [_i32p, _u8p, C.c_int, C.c_size_t, _i32p, _i32p, C.c_int, C.c_int,
                                                 _u8p, C.c_int, C.c_size_t, _f64p, _i32p, _i32p, C.POINTER(C.c_int)]
            L.plref_line_coords.restype = C.c_int
            L.plref_line_coords.argtypes = [C.c_double] * 4 + [_i32p, C.c_int]
            L.plref_normalize.argtypes = [_f64p]
            if hasattr(L, "plref_med_desc"):  # the reference build only (not the C++ drop-in harness)
                L.plref_med_desc.restype = None
                L.plref_med_desc.argtypes = [C.c_int, _u8p, C.c_size_t, _f64p, _i32p, C.c_int, _i32p, _f64p]
            self._lib = L
        return self._lib

    def med_desc(self, desc, dirs, obs_start, is_line=False):
        """PLSLAM::MapPoint / MapLine built observation by observation from the reference's own
        src/mapFeatures.cpp -> (med_idx, med_dir)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        obs_start, osp = _i32(obs_start)
        n_lm = len(obs_start) - 1
        dirs_c = None if dirs is None else np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        med_idx = np.empty(n_lm, np.int32)
        med_dir = np.zeros((n_lm, 3), np.float64)
        self.lib.plref_med_desc(int(bool(is_line)), desc.ctypes.data_as(_u8p), 32,
                                None if dirs_c is None else dirs_c.ctypes.data_as(_f64p), osp, n_lm,
                                med_idx.ctypes.data_as(_i32p), med_dir.ctypes.data_as(_f64p))
        return med_idx, med_dir

    def time_match(self, d1, d2, nnr, reps=100) -> float:
        """Median microseconds of StVO::match at the C++ signature level (reference build or GPU drop-in)."""
        p1, n1, s1 = _desc(np.ascontiguousarray(d1))
        p2, n2, s2 = _desc(np.ascontiguousarray(d2))
        self.lib.plref_time_match.restype = C.c_double
        self.lib.plref_time_match.argtypes = [_u8p, C.c_int, C.c_size_t, _u8p, C.c_int, C.c_size_t, C.c_float, C.c_int]
        return float(self.lib.plref_time_match(p1, n1, s1, p2, n2, s2, C.c_float(nnr), int(reps)))

    def time_match_grid(self, coords, d1, cell_start, cell_items, rows, cols, d2, win, dirs2=None, reps=100) -> float:
        """Median microseconds of StVO::matchGrid (points or, with dirs2, lines) for a prebuilt GridStructure."""
        is_lines = dirs2 is not None
        co, cop = _i32(np.ascontiguousarray(coords, np.int32).reshape(-1))
        p1, n1, s1 = _desc(np.ascontiguousarray(d1))
        p2, n2, s2 = _desc(np.ascontiguousarray(d2))
        cs, csp = _i32(cell_start)
        ci, cip = _i32(cell_items if len(cell_items) else np.zeros(1, np.int32))
        w, wp = _i32(win)
        dr = None if dirs2 is None else np.ascontiguousarray(dirs2, np.float64)
        self.lib.plref_time_match_grid.restype = C.c_double
        self.lib.plref_time_match_grid.argtypes = [C.c_int, _i32p, _u8p, C.c_int, C.c_size_t, _i32p, _i32p, C.c_int, C.c_int, _u8p, C.c_int,
                                                   C.c_size_t, _f64p, _i32p, C.c_int]
        return float(self.lib.plref_time_match_grid(int(is_lines), cop, p1, n1, s1, csp, cip, int(rows), int(cols), p2, n2, s2,
                                                    None if dr is None else dr.ctypes.data_as(_f64p), wp, int(reps)))

    def gpu_frame(self, pts, lines, t_pts, t_lines, rows, cols, nnr, stale=None, reps=0):
        """libstvo_gpu.so only: ONE frame through StVO::GpuFrame (the product's one-launch frame behind the StVO:: types).
        pts = (xy, d1, cell_start, cell_items, d2, win); lines = (xyxy, d1, cell_start, cell_items, d2, dirs2, win);
        t_pts / t_lines = (desc_prev, desc_curr).  stale: optional four in/out vectors.  Returns (counts[4], [m_sp, m_sl,
        m_tp, m_tl], median microseconds or None)."""
        c32 = lambda a: np.ascontiguousarray(a, np.int32).reshape(-1)  # noqa: E731
        d8 = lambda a: np.ascontiguousarray(a, np.uint8).reshape(-1, 32)  # noqa: E731
        xy, pd1, pcs, pci, pd2, pwin = pts
        xyxy, ld1, lcs, lci, ld2, dirs2, lwin = lines
        arrs = dict(xy=c32(xy), pd1=d8(pd1), pcs=c32(pcs), pci=c32(pci if len(pci) else np.zeros(1, np.int32)), pd2=d8(pd2), pwin=c32(pwin),
                    xyxy=c32(xyxy), ld1=d8(ld1), lcs=c32(lcs), lci=c32(lci if len(lci) else np.zeros(1, np.int32)), ld2=d8(ld2),
                    dirs=np.ascontiguousarray(dirs2, np.float64).reshape(-1), lwin=c32(lwin),
                    tp1=d8(t_pts[0]), tp2=d8(t_pts[1]), tl1=d8(t_lines[0]), tl2=d8(t_lines[1]))
        sizes = [len(arrs["pd1"]), len(arrs["ld1"]), len(arrs["tp1"]), len(arrs["tl1"])]
        m = [np.full(n, -1, np.int32) if stale is None else np.array(stale[k], np.int32) for k, n in enumerate(sizes)]
        counts = np.zeros(4, np.int32)
        med = C.c_double(0.0)
        u8 = lambda a: a.ctypes.data_as(_u8p)  # noqa: E731
        i32 = lambda a: a.ctypes.data_as(_i32p)  # noqa: E731
        fn = self.lib.plref_gpu_frame
        fn.restype = C.c_int
        fn.argtypes = [_i32p, _u8p, C.c_int, _i32p, _i32p, _u8p, C.c_int, _i32p, _i32p, _u8p, C.c_int, _i32p, _i32p, _u8p, C.c_int, _f64p,
                       _i32p, C.c_int, C.c_int, _u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_float, _i32p, _i32p,
                       _i32p, _i32p, _i32p, C.c_int, C.POINTER(C.c_double)]
        st = fn(i32(arrs["xy"]), u8(arrs["pd1"]), sizes[0], i32(arrs["pcs"]), i32(arrs["pci"]), u8(arrs["pd2"]), len(arrs["pd2"]),
                i32(arrs["pwin"]), i32(arrs["xyxy"]), u8(arrs["ld1"]), sizes[1], i32(arrs["lcs"]), i32(arrs["lci"]), u8(arrs["ld2"]),
                len(arrs["ld2"]), arrs["dirs"].ctypes.data_as(_f64p), i32(arrs["lwin"]), int(rows), int(cols), u8(arrs["tp1"]), sizes[2],
                u8(arrs["tp2"]), len(arrs["tp2"]), u8(arrs["tl1"]), sizes[3], u8(arrs["tl2"]), len(arrs["tl2"]), C.c_float(nnr),
                i32(m[0]), i32(m[1]), i32(m[2]), i32(m[3]), i32(counts), int(reps), C.byref(med))
        if st != 0:
            raise RuntimeError("plref_gpu_frame failed")
        return counts, m, (float(med.value) if reps > 0 else None)

    def set_threads(self, n: int):
        self.lib.plref_set_threads(int(n))

    def set_config(self, best_lr=True, lr_parallel=True, min_ratio_12p=0.9, line_sim_th=0.75):
        self.lib.plref_set_config(int(bool(best_lr)), int(bool(lr_parallel)), float(min_ratio_12p),
                                  float(line_sim_th))

    def distance(self, a, b) -> int:
        a = np.ascontiguousarray(a, np.uint8).reshape(32)
        b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return self.lib.plref_distance(a.ctypes.data_as(_u8p), b.ctypes.data_as(_u8p))

    def _nnr_like(self, fn, d1, d2, nnr, m12):
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = C.c_int(0)
        st = fn(p1, n1, s1, p2, n2, s2, C.c_float(nnr), m12.ctypes.data_as(_i32p), C.byref(n))
        if st != 0:
            raise RuntimeError("reference threw std::runtime_error")
        return n.value, m12

    def match_nnr(self, d1, d2, nnr, m12=None):
        return self._nnr_like(self.lib.plref_match_nnr, d1, d2, nnr, m12)

    def match(self, d1, d2, nnr, best_lr=True, lr_parallel=True, m12=None):
        self.set_config(best_lr=best_lr, lr_parallel=lr_parallel)
        return self._nnr_like(self.lib.plref_match, d1, d2, nnr, m12)

    def match_grid_points(self, xy, d1, cell_start, cell_items, rows, cols, d2, win, ratio, best_lr=True, m12=None):
        self.set_config(best_lr=best_lr, min_ratio_12p=ratio)
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        xy, xyp = _i32(xy)
        cs, csp = _i32(cell_start)
        ci, cip = _i32(cell_items if len(cell_items) else np.zeros(1, np.int32))
        w, wp = _i32(win)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = C.c_int(0)
        st = self.lib.plref_match_grid_points(xyp, p1, n1, s1, csp, cip, rows, cols, p2, n2, s2, wp,
                                              m12.ctypes.data_as(_i32p), C.byref(n))
        if st != 0:
            raise RuntimeError("reference threw std::runtime_error")
        return n.value, m12

    def match_grid_lines(self, xyxy, d1, cell_start, cell_items, rows, cols, d2, dirs2, line_sim_th, win, ratio,
                         best_lr=True, m12=None):
        self.set_config(best_lr=best_lr, min_ratio_12p=ratio, line_sim_th=line_sim_th)
        p1, n1, s1 = _desc(d1)
        p2, n2, s2 = _desc(d2)
        xy, xyp = _i32(xyxy)
        cs, csp = _i32(cell_start)
        ci, cip = _i32(cell_items if len(cell_items) else np.zeros(1, np.int32))
        w, wp = _i32(win)
        dirs2 = np.ascontiguousarray(dirs2, np.float64)
        m12 = np.full(n1, -1, np.int32) if m12 is None else np.array(m12, np.int32)
        n = C.c_int(0)
        st = self.lib.plref_match_grid_lines(xyp, p1, n1, s1, csp, cip, rows, cols, p2, n2, s2,
                                             dirs2.ctypes.data_as(_f64p), wp, m12.ctypes.data_as(_i32p),
                                             C.byref(n))
        if st != 0:
            raise RuntimeError("reference threw std::runtime_error")
        return n.value, m12

    def line_coords(self, x1, y1, x2, y2, max_cells=4096) -> np.ndarray:
        cells = np.empty((max_cells, 2), np.int32)
        n = self.lib.plref_line_coords(x1, y1, x2, y2, cells.ctypes.data_as(_i32p), max_cells)
        assert n <= max_cells
        return cells[:n].copy()

    def normalize(self, v) -> np.ndarray:
        v = np.array(v, np.float64)
        self.lib.plref_normalize(v.ctypes.data_as(_f64p))
        return v


class _RefStereo:
    """The reference's own stereo drivers and gates: stvo-pl/src/{stereoFrame,stereoFeatures,pinholeStereoCamera}.cpp
    (+ matching / gridStructure / lineIterator) compiled unmodified into ``_ref/libplref_stereo.so``
    (oracle/Makefile, stand-in headers under oracle/shim_stereo/).  Outputs use the layouts of ``oracle.port``'s
    ``stereo_points`` / ``stereo_lines``; the left index of a kept feature travels through KeyPoint/KeyLine::octave
    (StereoFrame copies it into PointFeature/LineFeature::level, stereoFrame.cpp:176,394)."""

    def __init__(self, libname: str = "libplref_stereo.so"):
        self._lib = None
        self._libname = libname

    def available(self) -> bool:
        try:
            return self.lib is not None
        except (FileNotFoundError, OSError, RuntimeError):
            return False

    @property
    def lib(self):
        if self._lib is None:
            L = _load(self._libname)
            L.plref_set_config.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
            L.plref_set_stereo_config.argtypes = [C.c_int] + [C.c_double] * 7
            L.plref_stereo_points.restype = C.c_int
            L.plref_stereo_points.argtypes = [_f32p, _i32p, _u8p, C.c_int, _f32p, _u8p, C.c_int, C.c_int, C.c_int, _f64p,
                                              C.c_int, _f64p, _f64p, _f64p, _i32p, _i32p, _f64p, _u8p, _i32p]
            L.plref_stereo_lines.restype = C.c_int
            L.plref_stereo_lines.argtypes = [_f32p, _f32p, _i32p, _u8p, C.c_int, _f32p, _u8p, C.c_int, C.c_int, C.c_int,
                                             _f64p, C.c_int] + [_f64p] * 7 + [_i32p, _i32p, _f64p, _u8p, _i32p]
            L.plref_filter_line_disparity.argtypes = [C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p]
            L.plref_filter_disparity_pair.argtypes = [C.c_int, _f64p]
            L.plref_line_overlap_stereo.argtypes = [C.c_int, _f64p, _f64p]
            L.plref_line_overlap.argtypes = [C.c_int, _f64p, _f64p]
            L.plref_back_projection.argtypes = [C.c_int, C.c_int, _f64p, C.c_int, _f64p, _f64p]
            self._lib = L
        return self._lib

    def configure(self, best_lr=True, ratio=0.9, line_sim_th=0.75, matching_s_ws=10, max_dist_epip=1.0, min_disp=1.0,
                  line_horiz_th=0.1, stereo_overlap_th=0.75, ls_min_disp_ratio=0.7, orb_scale_factor=1.2, lsd_scale=1.2):
        self.lib.plref_set_config(int(bool(best_lr)), 1, float(ratio), float(line_sim_th))
        self.lib.plref_set_stereo_config(int(matching_s_ws), float(max_dist_epip), float(min_disp), float(line_horiz_th),
                                         float(stereo_overlap_th), float(ls_min_disp_ratio), float(orb_scale_factor),
                                         float(lsd_scale))

    def stereo_points(self, kp_l, d_l, kp_r, d_r, img_w, img_h, cam, initial=False, octave=None, **cfg):
        """StereoFrame::matchStereoPoints -> dict(kept_i1, disp, P, pl, idx, level, sigma2, desc)."""
        self.configure(**cfg)
        kp_l = np.ascontiguousarray(kp_l, np.float32).reshape(-1, 2)
        kp_r = np.ascontiguousarray(kp_r, np.float32).reshape(-1, 2)
        d_l = np.ascontiguousarray(d_l, np.uint8).reshape(-1, 32)
        d_r = np.ascontiguousarray(d_r, np.uint8).reshape(-1, 32)
        n_l, n_r, cap = len(kp_l), len(kp_r), max(1, len(kp_l))
        cam = np.ascontiguousarray(cam, np.float64)
        tag = octave is None
        octv = np.arange(n_l, dtype=np.int32) if tag else np.ascontiguousarray(octave, np.int32)
        pl, disp, P = np.zeros((cap, 2)), np.zeros(cap), np.zeros((cap, 3))
        idx, level, sigma2 = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
        desc, nd = np.zeros((cap, 32), np.uint8), C.c_int32(0)
        n = self.lib.plref_stereo_points(kp_l.ctypes.data_as(_f32p), octv.ctypes.data_as(_i32p), d_l.ctypes.data_as(_u8p),
                                         n_l, kp_r.ctypes.data_as(_f32p), d_r.ctypes.data_as(_u8p), n_r, int(img_w),
                                         int(img_h), cam.ctypes.data_as(_f64p), int(bool(initial)),
                                         pl.ctypes.data_as(_f64p), disp.ctypes.data_as(_f64p), P.ctypes.data_as(_f64p),
                                         idx.ctypes.data_as(_i32p), level.ctypes.data_as(_i32p),
                                         sigma2.ctypes.data_as(_f64p), desc.ctypes.data_as(_u8p), C.byref(nd))
        n = max(n, 0)
        out = dict(disp=disp[:n].copy(), P=P[:n].copy(), pl=pl[:n].copy(), idx=idx[:n].copy(), level=level[:n].copy(),
                   sigma2=sigma2[:n].copy(), desc=desc[:nd.value].copy())
        if tag:
            out["kept_i1"] = level[:n].copy()
        return out

    def stereo_lines(self, ln_l, d_l, ln_r, d_r, img_w, img_h, cam, initial=False, octave=None, angle=None, **cfg):
        """StereoFrame::matchStereoLines -> dict(kept_i1, disp_se, sP, eP, le, spl, epl, angle, idx, level, sigma2, desc)."""
        self.configure(**cfg)
        ln_l = np.ascontiguousarray(ln_l, np.float32).reshape(-1, 4)
        ln_r = np.ascontiguousarray(ln_r, np.float32).reshape(-1, 4)
        d_l = np.ascontiguousarray(d_l, np.uint8).reshape(-1, 32)
        d_r = np.ascontiguousarray(d_r, np.uint8).reshape(-1, 32)
        n_l, n_r, cap = len(ln_l), len(ln_r), max(1, len(ln_l))
        cam = np.ascontiguousarray(cam, np.float64)
        tag = octave is None
        octv = np.arange(n_l, dtype=np.int32) if tag else np.ascontiguousarray(octave, np.int32)
        ang = np.zeros(n_l, np.float32) if angle is None else np.ascontiguousarray(angle, np.float32)
        spl, epl, dse = np.zeros((cap, 2)), np.zeros((cap, 2)), np.zeros((cap, 2))
        sP, eP, le = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3))
        ang_o, idx, level, sigma2 = np.zeros(cap), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
        desc, nd = np.zeros((cap, 32), np.uint8), C.c_int32(0)
        f64 = lambda a: a.ctypes.data_as(_f64p)  # noqa: E731
        n = self.lib.plref_stereo_lines(ln_l.ctypes.data_as(_f32p), ang.ctypes.data_as(_f32p), octv.ctypes.data_as(_i32p),
                                        d_l.ctypes.data_as(_u8p), n_l, ln_r.ctypes.data_as(_f32p),
                                        d_r.ctypes.data_as(_u8p), n_r, int(img_w), int(img_h), f64(cam), int(bool(initial)),
                                        f64(spl), f64(epl), f64(dse), f64(sP), f64(eP), f64(le), f64(ang_o),
                                        idx.ctypes.data_as(_i32p), level.ctypes.data_as(_i32p), f64(sigma2),
                                        desc.ctypes.data_as(_u8p), C.byref(nd))
        n = max(n, 0)
        out = dict(disp_se=dse[:n].copy(), sP=sP[:n].copy(), eP=eP[:n].copy(), le=le[:n].copy(), spl=spl[:n].copy(),
                   epl=epl[:n].copy(), angle=ang_o[:n].copy(), idx=idx[:n].copy(), level=level[:n].copy(),
                   sigma2=sigma2[:n].copy(), desc=desc[:nd.value].copy())
        if tag:
            out["kept_i1"] = level[:n].copy()
        return out

    def filter_line_disparity(self, spl, epl, spr, epr, ls_min_disp_ratio=0.7):
        self.configure(ls_min_disp_ratio=ls_min_disp_ratio)
        a = [np.ascontiguousarray(x, np.float64).reshape(-1, 2) for x in (spl, epl, spr, epr)]
        out = np.zeros((len(a[0]), 2))
        self.lib.plref_filter_line_disparity(len(a[0]), *[x.ctypes.data_as(_f64p) for x in a], out.ctypes.data_as(_f64p))
        return out

    def line_overlap_stereo(self, v4, line_horiz_th=0.1):
        self.configure(line_horiz_th=line_horiz_th)
        v4 = np.ascontiguousarray(v4, np.float64).reshape(-1, 4)
        out = np.zeros(len(v4))
        self.lib.plref_line_overlap_stereo(len(v4), v4.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
        return out

    def line_overlap(self, v8):
        v8 = np.ascontiguousarray(v8, np.float64).reshape(-1, 8)
        out = np.zeros(len(v8))
        self.lib.plref_line_overlap(len(v8), v8.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
        return out

    def back_projection(self, cam, uvd, img_w=752, img_h=480):
        cam = np.ascontiguousarray(cam, np.float64)
        uvd = np.ascontiguousarray(uvd, np.float64).reshape(-1, 3)
        out = np.zeros((len(uvd), 3))
        self.lib.plref_back_projection(int(img_w), int(img_h), cam.ctypes.data_as(_f64p), len(uvd),
                                       uvd.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
        return out


class _MapDropIn:
    """The product's C++ drop-in for src/mapFeatures.cpp (pl_inertial_slam_b200/csrc/map_features_gpu.cpp) behind the
    harness of oracle/shim/ref_map_capi.cpp -- the thing under test in tests/test_cxx_dropin.py, not a checker."""

    def __init__(self, libname="libmapfeatures_gpu.so"):
        self._libname, self._lib = libname, None

    def available(self) -> bool:
        return os.path.exists(os.path.join(_REF_DIR, self._libname))

    @property
    def lib(self):
        if self._lib is None:
            L = C.CDLL(os.path.join(_REF_DIR, self._libname))
            for f in ("plref_med_desc", "plref_med_desc_batch"):
                getattr(L, f).restype = None
                getattr(L, f).argtypes = [C.c_int, _u8p, C.c_size_t, _f64p, _i32p, C.c_int, _i32p, _f64p]
            self._lib = L
        return self._lib

    def med_desc(self, desc, dirs, obs_start, is_line=False, batch=False):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        obs_start, osp = _i32(obs_start)
        n_lm = len(obs_start) - 1
        dirs_c = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        med_idx = np.empty(n_lm, np.int32)
        med_dir = np.zeros((n_lm, 3), np.float64)
        fn = self.lib.plref_med_desc_batch if batch else self.lib.plref_med_desc
        fn(int(bool(is_line)), desc.ctypes.data_as(_u8p), 32, dirs_c.ctypes.data_as(_f64p), osp, n_lm,
           med_idx.ctypes.data_as(_i32p), med_dir.ctypes.data_as(_f64p))
        return med_idx, med_dir


port = _Port()
ref = _Ref()
ref_stereo = _RefStereo()
map_gpu = _MapDropIn()
stvo_gpu = _Ref("libstvo_gpu.so")  # the thing under test in tests/test_cxx_dropin.py, not a checker


_u32p = C.POINTER(C.c_uint32)


class FlatVocabulary:
    """A DBoW2 vocabulary tree as flat arrays (the layout of plm_voc_create, include/plmatch.h): node 0 is
    the root; children of node i are child_ids[child_start[i] : child_start[i+1]] in the reference's vector
    order; node_word[i] >= 0 marks a leaf (word id)."""

    def __init__(self, child_start, child_ids, node_desc, node_weight, node_word, k, L, weighting=0, scoring=0):
        self.child_start = np.ascontiguousarray(child_start, np.int32)
        self.child_ids = np.ascontiguousarray(child_ids, np.int32)
        self.node_desc = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        self.node_weight = np.ascontiguousarray(node_weight, np.float64)
        self.node_word = np.ascontiguousarray(node_word, np.int32)
        self.k, self.L, self.weighting, self.scoring = int(k), int(L), int(weighting), int(scoring)

    @property
    def n_nodes(self) -> int:
        return len(self.node_word)

    @property
    def n_words(self) -> int:
        return int(self.node_word.max()) + 1 if self.n_nodes else 0

    def arrays(self):
        return dict(child_start=self.child_start, child_ids=self.child_ids, node_desc=self.node_desc,
                    node_weight=self.node_weight, node_word=self.node_word,
                    meta=np.array([self.k, self.L, self.weighting, self.scoring], np.int32))

    @staticmethod
    def from_arrays(z, prefix=""):
        k, L, w, s = (int(x) for x in z[prefix + "meta"])
        return FlatVocabulary(z[prefix + "child_start"], z[prefix + "child_ids"], z[prefix + "node_desc"],
                              z[prefix + "node_weight"], z[prefix + "node_word"], k, L, w, s)


class _RefDbow:
    """The reference's vendored DBoW2 (3rdparty/DBoW2, compiled unmodified -> _ref/libplref_dbow.so):
    Vocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB> (include/mapHandler.h:70).  With
    libname="libdbow_gpu.so": the same harness over the product's drop-in vocabulary type PLM::GpuVocabulary
    (pl_inertial_slam_b200/csrc/dbow_vocabulary_gpu.h) -- the thing under test, not a checker."""

    def __init__(self, libname="libplref_dbow.so"):
        self._lib = None
        self._libname = libname

    def available(self) -> bool:
        if self._libname != "libplref_dbow.so":
            return os.path.exists(os.path.join(_REF_DIR, self._libname))
        try:
            return self.lib is not None
        except (FileNotFoundError, OSError, RuntimeError):
            return False

    @property
    def lib(self):
        if self._lib is None:
            if self._libname == "libplref_dbow.so":
                L = _load(self._libname)
            else:
                L = C.CDLL(os.path.join(_REF_DIR, self._libname))
            L.plref_voc_create.restype = C.c_void_p
            L.plref_voc_create.argtypes = [_u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            L.plref_voc_from_flat.restype = C.c_void_p
            L.plref_voc_from_flat.argtypes = [C.c_int, _i32p, _i32p, _u8p, _f64p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int]
            L.plref_voc_destroy.argtypes = [C.c_void_p]
            for f in ("plref_voc_n_nodes", "plref_voc_n_children", "plref_voc_n_words"):
                getattr(L, f).restype = C.c_int
                getattr(L, f).argtypes = [C.c_void_p]
            L.plref_voc_export.argtypes = [C.c_void_p, _i32p, _i32p, _u8p, _f64p, _i32p]
            L.plref_voc_transform.restype = C.c_int
            L.plref_voc_transform.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_size_t, _u32p, _f64p]
            L.plref_voc_score.restype = C.c_double
            L.plref_voc_score.argtypes = [C.c_void_p, _u32p, _f64p, C.c_int, _u32p, _f64p, C.c_int]
            if hasattr(L, "plref_voc_score_all"):
                L.plref_voc_score_all.restype = None
                L.plref_voc_score_all.argtypes = [C.c_void_p, _u32p, _f64p, C.c_int, _u32p, _f64p, _i32p, C.c_int, _f64p]
            self._lib = L
        return self._lib

    def create(self, desc, set_start, k=10, L=3, weighting=0, scoring=0, seed=1):
        """Vocabulary::create on training sets -> opaque handle (destroy() it)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        ss, ssp = _i32(set_start)
        return self.lib.plref_voc_create(desc.ctypes.data_as(_u8p), ssp, len(ss) - 1, k, L, weighting, scoring, seed)

    def from_flat(self, fv: FlatVocabulary):
        return self.lib.plref_voc_from_flat(fv.n_nodes, fv.child_start.ctypes.data_as(_i32p),
                                            fv.child_ids.ctypes.data_as(_i32p), fv.node_desc.ctypes.data_as(_u8p),
                                            fv.node_weight.ctypes.data_as(_f64p), fv.node_word.ctypes.data_as(_i32p),
                                            fv.k, fv.L, fv.weighting, fv.scoring)

    def destroy(self, h):
        self.lib.plref_voc_destroy(h)

    def export(self, h, k, L, weighting=0, scoring=0) -> FlatVocabulary:
        n, nc = self.lib.plref_voc_n_nodes(h), self.lib.plref_voc_n_children(h)
        cs, ci = np.zeros(n + 1, np.int32), np.zeros(max(nc, 1), np.int32)
        nd, nw, wd = np.zeros((n, 32), np.uint8), np.zeros(n, np.float64), np.zeros(n, np.int32)
        self.lib.plref_voc_export(h, cs.ctypes.data_as(_i32p), ci.ctypes.data_as(_i32p), nd.ctypes.data_as(_u8p),
                                  nw.ctypes.data_as(_f64p), wd.ctypes.data_as(_i32p))
        return FlatVocabulary(cs, ci[:nc], nd, nw, wd, k, L, weighting, scoring)

    def transform(self, h, desc):
        """Vocabulary::transform(features, BowVector) -> (word ids uint32, values float64), word order."""
        p, n, s = _desc(np.ascontiguousarray(desc, np.uint8).reshape(-1, 32))
        ids, vals = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
        m = self.lib.plref_voc_transform(h, p, n, s, ids.ctypes.data_as(_u32p), vals.ctypes.data_as(_f64p))
        return ids[:m].copy(), vals[:m].copy()

    def score(self, h, v1, v2) -> float:
        i1, x1 = np.ascontiguousarray(v1[0], np.uint32), np.ascontiguousarray(v1[1], np.float64)
        i2, x2 = np.ascontiguousarray(v2[0], np.uint32), np.ascontiguousarray(v2[1], np.float64)
        return self.lib.plref_voc_score(h, i1.ctypes.data_as(_u32p), x1.ctypes.data_as(_f64p), len(i1),
                                        i2.ctypes.data_as(_u32p), x2.ctypes.data_as(_f64p), len(i2))


    def score_all(self, h, q, db):
        """GpuVocabulary::scoreAll (drop-in library only)."""
        qi, qv = np.ascontiguousarray(q[0], np.uint32), np.ascontiguousarray(q[1], np.float64)
        lens = np.array([len(v[0]) for v in db], np.int32)
        di = np.ascontiguousarray(np.concatenate([np.asarray(v[0], np.uint32) for v in db]) if len(db) else np.zeros(1, np.uint32))
        dv = np.ascontiguousarray(np.concatenate([np.asarray(v[1], np.float64) for v in db]) if len(db) else np.zeros(1))
        out = np.zeros(len(db), np.float64)
        self.lib.plref_voc_score_all(h, qi.ctypes.data_as(_u32p), qv.ctypes.data_as(_f64p), len(qi), di.ctypes.data_as(_u32p),
                                     dv.ctypes.data_as(_f64p), lens.ctypes.data_as(_i32p), len(db), out.ctypes.data_as(_f64p))
        return out


ref_dbow = _RefDbow()
dbow_gpu = _RefDbow("libdbow_gpu.so")  # the thing under test in tests/test_cxx_dropin.py
