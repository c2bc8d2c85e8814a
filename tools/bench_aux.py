"""Side kernels only (landmark medoids, BoW transform / score): python tools/bench_aux.py [--cpu]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench_extras  # noqa: E402
from pl_inertial_slam_b200.matching import Context  # noqa: E402

if __name__ == "__main__":
    cpu = "--cpu" in sys.argv
    torch.cuda.set_device(0)
    ctx = Context()
    out = {"map_landmarks": bench_extras.map_landmarks(ctx, cpu=cpu), "bow_scoring": bench_extras.bow_scoring(ctx, cpu=cpu)}
    print(json.dumps(out))
