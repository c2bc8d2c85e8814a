"""Config 3 through the device-resident frame pipeline: python tools/bench_frames.py [n_frames]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extras  # noqa: E402
from pl_inertial_slam_b200 import matching as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = M.Context(0)
res = bench_extras.replay_pipeline(ctx, n)
if "--cpu" in sys.argv:
    res["cpu_baseline"] = bench_extras.replay_pipeline_cpu_baseline(min(n, 200))
print(json.dumps(res))
