"""CPU stand-in for pl_inertial_slam_b200.database.DeviceOps, built on the oracle, so that the
multi-rank ORCHESTRATION (sharding, exchanges, merges) can run under gloo without a GPU.
Test infrastructure only."""
import numpy as np
import torch

import oracle

port = oracle.port
ABSENT = np.uint64(0xFFFFFFFFFFFFFFFF)


def _np(t):
    return t.detach().cpu().numpy()


class CpuOps:
    def knn2(self, q, db, idx_base=0, out=None):
        key = port.knn2_packed(_np(q), _np(db), idx_base=idx_base) if q.shape[0] else np.zeros((0, 2), np.uint64)
        return torch.from_numpy(key.view(np.int64).copy())

    def top2_merge(self, parts, out=None):
        k = _np(parts).view(np.uint64)                      # P x nq x 2
        flat = np.sort(np.transpose(k, (1, 0, 2)).reshape(k.shape[1], -1), axis=1)
        return torch.from_numpy(flat[:, :2].copy().view(np.int64))

    def nnr_accept(self, top2, nnr, m12, count):
        k = _np(top2).view(np.uint64)
        d0 = (k[:, 0] >> np.uint64(32)).astype(np.float32)
        d1 = (k[:, 1] >> np.uint64(32)).astype(np.float32)
        acc = (k[:, 1] != ABSENT) & (d0 < d1 * np.float32(nnr))
        idx = (k[:, 0] & np.uint64(0xFFFFFFFF)).astype(np.int32)
        m = m12.numpy()
        m[acc] = idx[acc]
        if count is not None:
            count += int(acc.sum())

    def cross_check(self, m12, i1_base, m21, count):
        m = m12.numpy()
        r = m21.numpy()
        culled = 0
        for i1 in range(len(m)):
            i2 = m[i1]
            if i2 >= 0 and (i2 >= len(r) or r[i2] != i1_base + i1):
                m[i1] = -1
                culled += 1
        count -= culled

    def _shard(self, coords, d1, i1_base, frame, win, ratio, th, best_lr, m12, seed):
        dirs = _np(frame.dirs2) if frame.dirs2 is not None else None
        return port.match_grid_shard(dirs is not None, _np(coords), _np(d1), i1_base, _np(frame.cell_start),
                                     _np(frame.cell_items), frame.rows, frame.cols, _np(frame.d2), dirs, th,
                                     np.asarray(win, np.int32), ratio, best_lr, m12,
                                     None if seed is None else _np(seed).view(np.uint16))

    def grid_colmin(self, coords, d1, i1_base, frame, win, ratio, th, best_lr):
        n1 = d1.shape[0]
        _, _, colmin, _ = self._shard(coords, d1, i1_base, frame, win, ratio, th, best_lr, np.full(n1, -1, np.int32), None)
        return torch.from_numpy(colmin.view(np.int16).copy())

    def grid_match(self, coords, d1, i1_base, frame, win, ratio, th, best_lr, m12, count, seed):
        n, m, _, key = self._shard(coords, d1, i1_base, frame, win, ratio, th, best_lr, m12.numpy(), seed)
        m12.copy_(torch.from_numpy(m))
        count += n
        return torch.from_numpy(key.view(np.int64).copy())

    def m21_from_keys(self, key):
        k = _np(key).view(np.uint64)
        m21 = np.where(k == ABSENT, -1, (k & np.uint64(0xFFFFFFFF)).astype(np.int64)).astype(np.int32)
        return torch.from_numpy(m21)
