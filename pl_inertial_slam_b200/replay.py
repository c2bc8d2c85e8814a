"""Batched replay (config 3): one launch per stage over a frame arena instead of the per-frame loop of
app/plslam_dataset.cpp:114-172.  Thin wrapper over plm_batch_* (include/plmatch.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from . import grid as G
from .matching import Context
from .synth import Replay


class MatchBatch:
    """A prepared batch: set_*() uploads arenas + job tables, run() launches, fetch() reads back."""

    def __init__(self, ctx: Optional[Context] = None):
        self.lib = L.load()
        self.ctx = ctx
        self._h = C.c_void_p()
        L.check(self.lib.plm_batch_create(ctx.handle if ctx else None, C.byref(self._h)), "plm_batch_create")
        self.n_jobs = 0
        self.n_m = 0

    @staticmethod
    def _vp(a) -> C.c_void_p:
        if a is None:
            return C.c_void_p(0)
        if hasattr(a, "data_ptr"):            # torch (pinned) tensor
            return C.c_void_p(a.data_ptr())
        return C.c_void_p(a.ctypes.data)

    def set_match(self, arena, jobs: np.ndarray, nnr: float, best_lr: bool, m12_arena) -> None:
        assert jobs.dtype == L.PAIR_JOB_DTYPE
        self.n_jobs, self.n_m = len(jobs), len(m12_arena)
        L.check(self.lib.plm_batch_set_match(self._h, self._vp(arena), len(arena), self._vp(jobs), len(jobs),
                                             C.c_float(nnr), int(bool(best_lr)), self._vp(m12_arena), len(m12_arena)),
                "plm_batch_set_match")

    def set_match_dev(self, arena_dev, jobs: np.ndarray, nnr: float, best_lr: bool, n_m: int, m12_arena=None) -> None:
        """plm_batch_set_match_dev: the descriptor arena is a resident torch CUDA tensor, used in place."""
        assert jobs.dtype == L.PAIR_JOB_DTYPE and arena_dev.is_cuda and arena_dev.is_contiguous()
        self.n_jobs, self.n_m = len(jobs), int(n_m)
        self._arena_ref = arena_dev           # keep the tensor alive as long as the batch uses it
        L.check(self.lib.plm_batch_set_match_dev(self._h, self._vp(arena_dev), int(arena_dev.shape[0]), self._vp(jobs), len(jobs),
                                                 C.c_float(nnr), int(bool(best_lr)), self._vp(m12_arena), int(n_m)),
                "plm_batch_set_match_dev")

    def set_match_grid(self, arena, coords, cell_start, cell_items, dirs2, rows: int, cols: int, jobs: np.ndarray,
                       ratio: float, line_sim_th: float, best_lr: bool, m12_arena) -> None:
        assert jobs.dtype == L.GRID_JOB_DTYPE
        self.n_jobs, self.n_m = len(jobs), len(m12_arena)
        n_dirs = 0 if dirs2 is None else len(dirs2)
        L.check(self.lib.plm_batch_set_match_grid(self._h, self._vp(arena), len(arena), self._vp(coords), len(coords),
                                                  self._vp(cell_start), len(cell_start), self._vp(cell_items),
                                                  len(cell_items), self._vp(dirs2), n_dirs, rows, cols, self._vp(jobs),
                                                  len(jobs), float(ratio), float(line_sim_th), int(bool(best_lr)),
                                                  self._vp(m12_arena), len(m12_arena)), "plm_batch_set_match_grid")

    def run(self) -> None:
        L.check(self.lib.plm_batch_run(self._h), "plm_batch_run")

    def fetch_counts(self) -> np.ndarray:
        """Only the per-job return values (the match vectors stay on the device)."""
        counts = np.empty(self.n_jobs, np.int32)
        L.check(self.lib.plm_batch_fetch(self._h, C.c_void_p(0), self._vp(counts)), "plm_batch_fetch")
        return counts

    def fetch(self, m12_out=None, counts_out=None) -> Tuple[np.ndarray, np.ndarray]:
        m12 = np.empty(self.n_m, np.int32) if m12_out is None else m12_out
        counts = np.empty(self.n_jobs, np.int32) if counts_out is None else counts_out
        L.check(self.lib.plm_batch_fetch(self._h, self._vp(m12), self._vp(counts)), "plm_batch_fetch")
        return m12, counts

    @property
    def h2d_bytes(self) -> int:
        return int(self.lib.plm_batch_h2d_bytes(self._h))

    @property
    def d2h_bytes(self) -> int:
        return int(self.lib.plm_batch_d2h_bytes(self._h))

    def close(self):
        if self._h:
            self.lib.plm_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stereo_grid_jobs(rp: Replay, matching_s_ws: int = 10) -> np.ndarray:
    """Per frame: matchGrid(points) (stereoFrame.cpp:157) and matchGrid(lines) (:356), left vs right."""
    F = rp.n_frames
    n_cells1 = G.GRID_ROWS * G.GRID_COLS + 1
    jobs = np.zeros(2 * F, L.GRID_JOB_DTYPE)
    p, l = jobs[0::2], jobs[1::2]
    p["off_coords"], p["off1"], p["off2"] = rp.off_cpts, rp.off_pl, rp.off_pr
    p["off_cell_start"] = np.arange(F) * 2 * n_cells1
    p["off_cell_items"], p["off_m"], p["n1"], p["n2"], p["is_lines"] = rp.off_items_p, rp.off_m_p, rp.n_pts, rp.n_pts, 0
    l["off_coords"], l["off1"], l["off2"] = rp.off_clines, rp.off_ll, rp.off_lr
    l["off_cell_start"] = np.arange(F) * 2 * n_cells1 + n_cells1
    l["off_cell_items"], l["off_dirs2"], l["off_m"] = rp.off_items_l, rp.off_dirs, rp.off_m_l
    l["n1"], l["n2"], l["is_lines"] = rp.n_lines, rp.n_lines, 1
    jobs["win"] = np.array([matching_s_ws, 0, 0, 0], np.int32)
    return jobs


def temporal_match_jobs(rp: Replay) -> np.ndarray:
    """Per frame f >= 1: match(prev left, curr left) for points and lines (stereoFrameHandler.cpp:168,191).
    Frame 0 has no predecessor; its two jobs are empty (n2 = 0 -> skipped)."""
    F = rp.n_frames
    jobs = np.zeros(2 * F, L.PAIR_JOB_DTYPE)
    p, l = jobs[0::2], jobs[1::2]
    p["off1"][1:], p["n1"][1:] = rp.off_pl[:-1], rp.n_pts[:-1]
    p["off2"][1:], p["n2"][1:] = rp.off_pl[1:], rp.n_pts[1:]
    l["off1"][1:], l["n1"][1:] = rp.off_ll[:-1], rp.n_lines[:-1]
    l["off2"][1:], l["n2"][1:] = rp.off_ll[1:], rp.n_lines[1:]
    # results of frame f's temporal jobs land in the slots of frame f-1's features (the query side)
    p["off_m"][1:] = rp.off_m_p[:-1]
    l["off_m"][1:] = rp.off_m_l[:-1]
    return jobs
