"""Per-call latency of the four per-frame matching calls (configs 1-2) through the host-buffer API."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import matching as M  # noqa: E402
from pl_inertial_slam_b200 import synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
ctx = M.Context(0)
prev, curr = synth.make_temporal_pair(synth.SEED0 + 2)
a = synth.stereo_points_grid_args(prev)
b = synth.stereo_lines_grid_args(prev)
ga = (a["cell_start"], a["cell_items"], a["rows"], a["cols"])
gb = (b["cell_start"], b["cell_items"], b["rows"], b["cols"])
k = synth.kf_points_grid_args(prev, curr)
gk = (k["cell_start"], k["cell_items"], k["rows"], k["cols"])
calls = {
    "stereo_matchGrid_points_600_w10": lambda: M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], [], ctx=ctx),
    "stereo_matchGrid_lines_200_w10": lambda: M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], [], ctx=ctx),
    "kf_matchGrid_points_600_w3": lambda: M.matchGrid(k["xy"], k["d1"], gk, k["d2"], k["win"], [], ctx=ctx),
    "temporal_match_points_600": lambda: M.match(prev.pdesc_l, curr.pdesc_l, 0.9, [], ctx=ctx),
    "temporal_match_lines_200": lambda: M.match(prev.ldesc_l, curr.ldesc_l, 0.9, [], ctx=ctx),
}
for name, fn in calls.items():
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    print(f"{name:36s} median {np.median(ts) * 1e6:8.1f} us   min {np.min(ts) * 1e6:8.1f} us")
