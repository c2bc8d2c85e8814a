"""First-contact probe on the B200: integer-pipe issue rates and brute-force kernel throughput."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import _lib as L  # noqa: E402
from pl_inertial_slam_b200.database import DeviceOps  # noqa: E402

out = {}
ops = DeviceOps(0)
popc, lop3 = ops.ctx.measure_int_peaks()
out["popc_gops"], out["lop3_gops"] = popc, lop3
print(f"POPC {popc:.1f} Gop/s  LOP3 {lop3:.1f} Gop/s", flush=True)

g = torch.Generator(device="cuda").manual_seed(1)
for (nq, n) in [(800, 2_000_000), (6400, 2_000_000), (6400, 16_000_000), (600, 600)]:
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
    db = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    res = {}
    outs = {}
    for variant in (2, 1, 0):
        L.load().plm_set_option(b"knn_variant", variant)
        o = ops.knn2(q, db)
        torch.cuda.synchronize()
        outs[variant] = o.clone()
        reps = 5 if nq * n < 3e10 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.knn2(q, db, out=o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        name = {0: "popc8", 1: "csa5", 2: "csa4_blocked"}[variant]
        res[name] = dict(ms=ms, gpairs_s=nq * n / ms * 1e-6)
        print(nq, n, name, f"{ms:.3f} ms  {nq * n / ms * 1e-6:.1f} Gpairs/s", flush=True)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "variants disagree"
    out[f"knn2_{nq}x{n}"] = res
    del q, db
L.load().plm_set_option(b"knn_variant", -1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
