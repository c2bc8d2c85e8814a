"""A/B of the frame pipeline's CTA widths (plm_set_option frames_threads_p / frames_threads_l): replay frames/s.
    python tools/ab_frames_threads.py [n_frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extras  # noqa: E402
from pl_inertial_slam_b200 import _lib, synth  # noqa: E402
from pl_inertial_slam_b200 import matching as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
lib = _lib.load()
ctx = M.Context(0)
rp = synth.make_replay(synth.SEED0 + 3, n)
for tp, tl in ((256, 128), (256, 256), (512, 256), (512, 128), (256, 256), (512, 256)):
    lib.plm_set_option(b"frames_threads_p", tp)
    lib.plm_set_option(b"frames_threads_l", tl)
    r = bench_extras.replay_pipeline(ctx, n, rp)
    print(f"points {tp} lines {tl}: device-resident {r['device_resident']['frames_per_s']:.0f} frames/s, e2e {r['e2e']['frames_per_s']:.0f}")
