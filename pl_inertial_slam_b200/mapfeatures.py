"""Host mirror of the reference's map landmarks (src/mapFeatures.cpp, include/mapFeatures.h:40-101) for
the one computation on the descriptor path: ``updateAverageDescDir`` -- the representative descriptor
(``med_desc``) and mean observation direction (``med_obs_dir``) of a MapPoint / MapLine.  ``med_desc``
rows are what map-to-keyframe matching reads (mapHandler.cpp:596-609, :698-714).

Same names and argument meaning as the reference; the arithmetic runs on the GPU through
``plm_med_desc`` / ``plm_dev_med_desc`` (include/plmatch.h).  The reference recomputes one landmark per
added observation; here ``update_average_desc_dir`` takes any number of landmarks in one launch, which
is how a keyframe insertion (hundreds of landmarks touched, mapHandler.cpp:144-283) should call it.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .matching import Context


def _ctx_handle(ctx: Optional[Context]):
    return ctx.handle if ctx is not None else None


def med_desc_batch(desc_obs: np.ndarray, obs_start: Sequence[int], dir_obs: Optional[np.ndarray] = None,
                   ctx: Optional[Context] = None):
    """plm_med_desc on host arrays -> (med_idx int32[n_lm], med_desc uint8[n_lm, 32], med_dir float64[n_lm, 3] | None).

    desc_obs: n_obs x 32 uint8 (rows may be strided); obs_start: n_lm + 1 offsets; dir_obs: n_obs x 3."""
    lib = L.load()
    desc_obs, dp, n_obs, step = L.desc_args(np.asarray(desc_obs, np.uint8).reshape(-1, 32)
                                            if not isinstance(desc_obs, np.ndarray) else desc_obs)
    obs_start = np.ascontiguousarray(obs_start, np.int32)
    n_lm = len(obs_start) - 1
    if n_lm < 0:
        raise ValueError("obs_start needs at least one entry")
    med_idx = np.empty(n_lm, np.int32)
    med = np.zeros((n_lm, 32), np.uint8)
    med_dir = None
    dirp = C.cast(None, L.f64p)
    mdp = C.cast(None, L.f64p)
    if dir_obs is not None:
        dir_obs = np.ascontiguousarray(dir_obs, np.float64).reshape(-1, 3)
        if len(dir_obs) != n_obs:
            raise ValueError("dir_obs and desc_obs differ in length")
        med_dir = np.zeros((n_lm, 3), np.float64)
        dirp, mdp = dir_obs.ctypes.data_as(L.f64p), med_dir.ctypes.data_as(L.f64p)
    L.check(lib.plm_med_desc(_ctx_handle(ctx), dp, n_obs, step, dirp, obs_start.ctypes.data_as(L.i32p), n_lm,
                             med_idx.ctypes.data_as(L.i32p), med.ctypes.data_as(L.u8p), mdp), "plm_med_desc")
    return med_idx, med, med_dir


def dev_med_desc(ctx: Context, desc_obs, obs_start, med_idx, med_desc=None, dir_obs=None, med_dir=None,
                 dst_rows=None) -> None:
    """plm_dev_med_desc on torch CUDA tensors (enqueued on the context's stream, no host sync).
    desc_obs uint8 [n_obs, 32]; obs_start int32 [n_lm + 1]; med_idx int32 [n_lm]; med_desc uint8 [*, 32];
    dst_rows int32 [n_lm] scatters landmark l to row dst_rows[l] of med_desc (a resident map shard)."""
    lib = L.load()
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
    n_lm = int(obs_start.numel()) - 1
    L.check(lib.plm_dev_med_desc(ctx.handle, ptr(desc_obs), int(desc_obs.shape[0]), ptr(dir_obs), ptr(obs_start), n_lm,
                                 ptr(med_idx), ptr(med_desc), ptr(dst_rows), ptr(med_dir)), "plm_dev_med_desc")


class _Landmark:
    """Common part of MapPoint / MapLine (mapFeatures.h:40-68, :70-101): the observation lists that
    updateAverageDescDir reads, and its two results."""

    def __init__(self, idx_: int, desc_: np.ndarray, kf_obs_: int, dir_: Sequence[float]):
        self.idx = idx_
        self.inlier = True
        self.desc_list: List[np.ndarray] = [np.ascontiguousarray(desc_, np.uint8).reshape(32)]
        self.kf_obs_list: List[int] = [kf_obs_]
        self.dir_list: List[np.ndarray] = [np.asarray(dir_, np.float64).reshape(3)]
        self.med_obs_dir = self.dir_list[0].copy()   # constructor, mapFeatures.cpp:37-38 / :106-107
        self.med_desc = self.desc_list[0]

    def _add(self, desc_, kf_obs_, dir_, ctx):
        self.desc_list.append(np.ascontiguousarray(desc_, np.uint8).reshape(32))
        self.kf_obs_list.append(kf_obs_)
        self.dir_list.append(np.asarray(dir_, np.float64).reshape(3))
        self.updateAverageDescDir(ctx=ctx)

    def updateAverageDescDir(self, ctx: Optional[Context] = None) -> None:
        """mapFeatures.cpp:51-93 / :121-163 for this landmark alone (one tiny launch; prefer the batch)."""
        update_average_desc_dir([self], ctx=ctx)


class MapPoint(_Landmark):
    """PLSLAM::MapPoint (mapFeatures.h:40-68)."""

    def __init__(self, idx_, point3D_, desc_, kf_obs_, obs_, dir_, sigma2_: float = 1.0):
        super().__init__(idx_, desc_, kf_obs_, dir_)
        self.point3D = np.asarray(point3D_, np.float64)
        self.obs_list = [np.asarray(obs_, np.float64)]
        self.sigma_list = [sigma2_]

    def addMapPointObservation(self, desc_, kf_obs_, obs_, dir_, sigma2_: float = 1.0, ctx: Optional[Context] = None,
                               defer: bool = False):
        """mapFeatures.cpp:42-49.  defer=True appends without recomputing (call update_average_desc_dir on
        the touched landmarks afterwards -- same final state, one launch)."""
        self.obs_list.append(np.asarray(obs_, np.float64))
        self.sigma_list.append(sigma2_)
        if defer:
            self.desc_list.append(np.ascontiguousarray(desc_, np.uint8).reshape(32))
            self.kf_obs_list.append(kf_obs_)
            self.dir_list.append(np.asarray(dir_, np.float64).reshape(3))
        else:
            self._add(desc_, kf_obs_, dir_, ctx)


class MapLine(_Landmark):
    """PLSLAM::MapLine (mapFeatures.h:70-101)."""

    def __init__(self, idx_, line3D_, desc_, kf_obs_, obs_, dir_, pts_, sigma2_: float = 1.0):
        super().__init__(idx_, desc_, kf_obs_, dir_)
        self.line3D = np.asarray(line3D_, np.float64)
        self.obs_list = [np.asarray(obs_, np.float64)]
        self.pts_list = [np.asarray(pts_, np.float64)]
        self.sigma_list = [sigma2_]

    def addMapLineObservation(self, desc_, kf_obs_, obs_, dir_, pts_, sigma2_: float = 1.0,
                              ctx: Optional[Context] = None, defer: bool = False):
        """mapFeatures.cpp:112-119."""
        self.obs_list.append(np.asarray(obs_, np.float64))
        self.pts_list.append(np.asarray(pts_, np.float64))
        self.sigma_list.append(sigma2_)
        if defer:
            self.desc_list.append(np.ascontiguousarray(desc_, np.uint8).reshape(32))
            self.kf_obs_list.append(kf_obs_)
            self.dir_list.append(np.asarray(dir_, np.float64).reshape(3))
        else:
            self._add(desc_, kf_obs_, dir_, ctx)


def update_average_desc_dir(landmarks: Iterable[_Landmark], ctx: Optional[Context] = None) -> np.ndarray:
    """updateAverageDescDir for every landmark of the iterable in ONE plm_med_desc call; sets med_desc /
    med_obs_dir on each and returns the winning list positions."""
    landmarks = list(landmarks)
    if not landmarks:
        return np.zeros(0, np.int32)
    counts = np.array([len(lm.desc_list) for lm in landmarks], np.int64)
    obs_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    desc = np.stack([d for lm in landmarks for d in lm.desc_list]) if counts.sum() else np.zeros((0, 32), np.uint8)
    dirs = np.stack([d for lm in landmarks for d in lm.dir_list]) if counts.sum() else np.zeros((0, 3))
    med_idx, _, med_dir = med_desc_batch(desc, obs_start, dirs, ctx=ctx)
    for lm, i, d in zip(landmarks, med_idx, med_dir):
        if i >= 0:
            lm.med_desc = lm.desc_list[int(i)]
            lm.med_obs_dir = d.copy()
    return med_idx
