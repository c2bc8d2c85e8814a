"""Small driver for ncu captures of the brute-force slice kernel: python tools/profile_knn2.py [variant] [nq] [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import _lib as L  # noqa: E402
from pl_inertial_slam_b200.database import DeviceOps  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else -1
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 6400
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
ops = DeviceOps(0)
L.load().plm_set_option(b"knn_variant", variant)
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
db = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
for _ in range(3):
    o = ops.knn2(q, db)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
o = ops.knn2(q, db)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"variant {variant} {nq}x{n}: {ms:.3f} ms {nq * n / ms * 1e-6:.1f} Gpairs/s")
