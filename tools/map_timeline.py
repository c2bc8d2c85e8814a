"""Debug (PLM_BUILD_DEFINES=-DPLM_TIMELINE builds only): phase stamps of the map-scale matchGrid kernels
(grid_rows_kernel pass 0 / pass 1) for config 4, mean over CTAs, in SM cycles since kernel entry."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import _lib as L  # noqa: E402
from pl_inertial_slam_b200 import grid as G  # noqa: E402
from pl_inertial_slam_b200 import synth  # noqa: E402
from pl_inertial_slam_b200.database import GridFrame, ShardedMap  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lines = "lines" in sys.argv
    sp = synth.make_stereo_pair(synth.SEED0 + 4)
    n_map = 200_000
    d1, xy = synth.make_map_points(synth.SEED0 + 4, n_map, sp)
    c = sp.kp_l.astype(np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * synth.INV_W, c[:, 1] * synth.INV_H)
    frame = GridFrame(torch.from_numpy(sp.pdesc_l).to(dev), torch.from_numpy(cs).to(dev), torch.from_numpy(ci).to(dev),
                      G.GRID_ROWS, G.GRID_COLS)
    smap = ShardedMap(n_map, torch.from_numpy(d1).to(dev), torch.from_numpy(xy).to(dev), device=0)
    win = np.array([3, 3, 3, 3], np.int32)
    lib = L.load()
    fn = lib.plm_debug_timeline
    fn.argtypes = [C.c_void_p]
    buf = np.zeros((128, 24), np.int64)
    acc = np.zeros((128, 24))
    runs = 50
    for r in range(runs + 5):
        smap.match_grid(frame, win, 0.9, 0.75, True)
        torch.cuda.synchronize()
        if r < 5:
            continue
        assert fn(buf.ctypes.data) == 0
        a = buf.astype(np.float64)
        rel = a.copy()
        rel[:, 1:6] -= a[:, 0:1]                       # pass 0 stamps relative to its entry
        for k in (16, 7, 8, 9, 10):
            rel[:, k] = a[:, k] - a[:, 15]             # pass 1 stamps relative to its entry
        acc += rel / runs
    names0 = {1: "staged", 2: "phaseA", 3: "emit", 4: "phaseB", 5: "end"}
    names1 = {16: "setup", 7: "thr-init", 8: "filter", 9: "rounds", 10: "end"}
    for label, ctas in (("cta 0", [0]), ("cta 1", [1]), ("cta 2-9", range(2, 10)), ("cta 10-126", range(10, 127))):
        row = acc[list(ctas)].mean(0)
        print(label, "pass0 (last block):", {v: int(row[k]) for k, v in names0.items()},
              "pass1 (last record):", {v: int(row[k]) for k, v in names1.items()})


if __name__ == "__main__":
    main()
