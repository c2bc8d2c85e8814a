"""Device-resident stereo-frame pipeline (plm_frames_* of include/plmatch.h).

Per frame, from raw features to tracked features without a host round trip:
StereoFrame::matchStereoPoints / matchStereoLines (stvo-pl/src/stereoFrame.cpp:131-184, :320-409)
followed by StereoFrameHandler::matchF2FPoints / matchF2FLines against the previous frame of the upload
(stvo-pl/src/stereoFrameHandler.cpp:158-207).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import _lib as L
from .matching import Context


@dataclass
class FrameConfig:
    """The Config / camera values the two drivers read (defaults: config.cpp:36-113, EuRoC-like camera)."""
    img_width: int = 752
    img_height: int = 480
    grid_rows: int = 48            # GRID_ROWS, stereoFrame.h:51
    grid_cols: int = 64            # GRID_COLS, stereoFrame.h:52
    matchingSWs: int = 10
    bestLRMatches: bool = True
    minRatio12P: float = 0.9
    minRatio12L: float = 0.9
    lineSimTh: float = 0.75
    maxDistEpip: float = 1.0
    minDisp: float = 1.0
    lineHorizTh: float = 0.1
    stereoOverlapTh: float = 0.75
    lsMinDispRatio: float = 0.7
    cam_b: float = 0.11
    cam_fx: float = 458.654
    cam_cx: float = 367.215
    cam_cy: float = 248.375

    @property
    def inv_width(self) -> float:
        return self.grid_cols / float(self.img_width)

    @property
    def inv_height(self) -> float:
        return self.grid_rows / float(self.img_height)

    @property
    def cam(self):
        return [self.cam_b, self.cam_fx, self.cam_cx, self.cam_cy]

    def to_c(self) -> L.FrameConfig:
        c = L.FrameConfig()
        c.inv_width, c.inv_height = self.inv_width, self.inv_height
        c.grid_rows, c.grid_cols = self.grid_rows, self.grid_cols
        c.matching_s_ws, c.best_lr = self.matchingSWs, int(self.bestLRMatches)
        c.min_ratio_12p, c.min_ratio_12l = self.minRatio12P, self.minRatio12L
        c.line_sim_th, c.max_dist_epip, c.min_disp = self.lineSimTh, self.maxDistEpip, self.minDisp
        c.line_horiz_th, c.stereo_overlap_th, c.ls_min_disp_ratio = self.lineHorizTh, self.stereoOverlapTh, self.lsMinDispRatio
        c.cam_b, c.cam_fx, c.cam_cx, c.cam_cy = self.cam_b, self.cam_fx, self.cam_cx, self.cam_cy
        return c


def _vp(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class FramePipeline:
    """upload() copies the arenas, run() launches both stages, fetch() reads the outputs back."""

    OUTPUTS = {"stereo_m12_p": ("P", 1, np.int32), "stereo_m12_l": ("L", 1, np.int32), "kept_p": ("P", 1, np.int32),
               "kept_l": ("L", 1, np.int32), "pt_disp": ("P", 1, np.float64), "pt_P": ("P", 3, np.float64),
               "ls_disp": ("L", 2, np.float64), "ls_sP": ("L", 3, np.float64), "ls_eP": ("L", 3, np.float64),
               "ls_le": ("L", 3, np.float64), "f2f_m12_p": ("P", 1, np.int32), "f2f_m12_l": ("L", 1, np.int32),
               "counts": ("F", 6, np.int32)}

    def __init__(self, ctx: Optional[Context] = None):
        self.lib = L.load()
        self.ctx = ctx
        self._h = C.c_void_p()
        L.check(self.lib.plm_frames_create(ctx.handle if ctx else None, C.byref(self._h)), "plm_frames_create")
        self.n_frames = self.NP = self.NL = 0
        self.off_p = self.off_l = None

    def upload(self, desc_arena, kp_arena, ln_arena, frames: np.ndarray, cfg: FrameConfig) -> None:
        """desc_arena n x 32 uint8, kp_arena k x 2 float32, ln_arena m x 4 float32 (numpy or pinned torch),
        frames = FRAME_REC_DTYPE records."""
        assert frames.dtype == L.FRAME_REC_DTYPE
        frames = np.ascontiguousarray(frames)
        c = cfg.to_c()
        L.check(self.lib.plm_frames_upload(self._h, _vp(desc_arena), len(desc_arena), _vp(kp_arena), len(kp_arena),
                                           _vp(ln_arena), len(ln_arena), _vp(frames), len(frames), C.byref(c)),
                "plm_frames_upload")
        self.set_layout(frames)

    def run(self) -> None:
        L.check(self.lib.plm_frames_run(self._h), "plm_frames_run")

    def set_layout(self, frames: np.ndarray) -> None:
        self.n_frames = len(frames)
        self.off_p = np.concatenate([[0], np.cumsum(frames["n_pl"].astype(np.int64))])
        self.off_l = np.concatenate([[0], np.cumsum(frames["n_ll"].astype(np.int64))])
        self.NP, self.NL = int(self.off_p[-1]), int(self.off_l[-1])

    def process(self, desc_arena, kp_arena, ln_arena, frames: np.ndarray, cfg: FrameConfig,
                out: Optional[Dict[str, np.ndarray]] = None, chunk_frames: int = 320) -> Dict[str, np.ndarray]:
        """upload + run + fetch as one chunk-pipelined call (plm_frames_process)."""
        assert frames.dtype == L.FRAME_REC_DTYPE
        frames = np.ascontiguousarray(frames)
        self.set_layout(frames)
        out = self.alloc_outputs() if out is None else out
        o = L.FramesOut()
        for name in self.OUTPUTS:
            setattr(o, name, _vp(out.get(name)))
        c = cfg.to_c()
        L.check(self.lib.plm_frames_process(self._h, _vp(desc_arena), len(desc_arena), _vp(kp_arena), len(kp_arena),
                                            _vp(ln_arena), len(ln_arena), _vp(frames), len(frames), C.byref(c),
                                            C.byref(o), int(chunk_frames)), "plm_frames_process")
        return out

    def alloc_outputs(self, names=None, pinned: bool = False) -> Dict[str, np.ndarray]:
        out = {}
        size = {"P": self.NP, "L": self.NL, "F": self.n_frames}
        for name in (names or self.OUTPUTS):
            kind, width, dt = self.OUTPUTS[name]
            shape = (size[kind],) if width == 1 else (size[kind], width)
            if pinned:
                import torch
                out[name] = torch.empty(shape, dtype=torch.int32 if dt == np.int32 else torch.float64).pin_memory()
            else:
                out[name] = np.empty(shape, dt)
        return out

    def fetch(self, out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        out = self.alloc_outputs() if out is None else out
        o = L.FramesOut()
        for name in self.OUTPUTS:
            setattr(o, name, _vp(out.get(name)))
        L.check(self.lib.plm_frames_fetch(self._h, C.byref(o)), "plm_frames_fetch")
        return out

    @property
    def h2d_bytes(self) -> int:
        return int(self.lib.plm_frames_h2d_bytes(self._h))

    @property
    def d2h_bytes(self) -> int:
        return int(self.lib.plm_frames_d2h_bytes(self._h))

    def close(self):
        if self._h:
            self.lib.plm_frames_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def replay_frame_records(rp) -> tuple:
    """(kp_arena, ln_arena, frame records) for a synth.Replay.  Keypoints and segments are laid out frame by frame
    ([left of frame f | right of frame f]), like the descriptor arena of the generator, so that the rows a chunk of
    consecutive frames needs are ONE contiguous slice of each arena (one host -> device copy per arena and chunk)."""
    F = rp.n_frames
    n_p, n_l = rp.n_pts.astype(np.int64), rp.n_lines.astype(np.int64)
    pbase = np.concatenate([[0], np.cumsum(n_p)])
    lbase = np.concatenate([[0], np.cumsum(n_l)])
    kp_arena = np.empty((2 * int(pbase[-1]), 2), np.float32)
    ln_arena = np.empty((2 * int(lbase[-1]), 4), np.float32)
    rec = np.zeros(F, L.FRAME_REC_DTYPE)
    rec["desc_pl"], rec["desc_pr"], rec["desc_ll"], rec["desc_lr"] = rp.off_pl, rp.off_pr, rp.off_ll, rp.off_lr
    rec["kp_l"], rec["kp_r"] = 2 * pbase[:-1], 2 * pbase[:-1] + n_p
    rec["ln_l"], rec["ln_r"] = 2 * lbase[:-1], 2 * lbase[:-1] + n_l
    # scatter [all left | all right] into the per-frame layout
    idx_p = np.repeat(pbase[:-1], n_p) if F else np.zeros(0, np.int64)            # first left row of every row's frame
    within_p = np.arange(int(pbase[-1])) - idx_p
    dst_p = 2 * idx_p + within_p
    kp_arena[dst_p] = np.asarray(rp.kp_l, np.float32).reshape(-1, 2)
    kp_arena[dst_p + np.repeat(n_p, n_p)] = np.asarray(rp.kp_r, np.float32).reshape(-1, 2)
    idx_l = np.repeat(lbase[:-1], n_l) if F else np.zeros(0, np.int64)
    within_l = np.arange(int(lbase[-1])) - idx_l
    dst_l = 2 * idx_l + within_l
    ln_arena[dst_l] = np.asarray(rp.ln_l, np.float32).reshape(-1, 4)
    ln_arena[dst_l + np.repeat(n_l, n_l)] = np.asarray(rp.ln_r, np.float32).reshape(-1, 4)
    rec["n_pl"] = rec["n_pr"] = n_p
    rec["n_ll"] = rec["n_lr"] = n_l
    return kp_arena, ln_arena, rec
