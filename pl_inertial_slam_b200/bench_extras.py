"""Side metrics of bench.py for the frame-scale configs of BASELINE.json (1-3): per-call latency of the
stereo / temporal matchers through the host-buffer C ABI, and replay throughput (matched stereo
frames/s) of the batched path, device-resident and end to end.  The CPU figures next to them come
from the oracle builds (reference sources where available) on this box's host cores."""
from __future__ import annotations

import os
import time

import numpy as np

from . import matching as M
from . import replay, synth


def _median_ms(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts) * 1e3)


def frame_latency(ctx) -> dict:
    """Config 1 + 2: one EuRoC-shaped stereo pair (600 ORB + 200 LBD) and its temporal match, host
    buffers in, host match vectors out, one call each (wall clock, median of 30)."""
    import oracle
    prev, curr = synth.make_temporal_pair(synth.SEED0 + 2)
    a = synth.stereo_points_grid_args(prev)
    b = synth.stereo_lines_grid_args(prev)
    ga = (a["cell_start"], a["cell_items"], a["rows"], a["cols"])
    gb = (b["cell_start"], b["cell_items"], b["rows"], b["cols"])
    M.Config.minRatio12P = 0.9
    out = {}
    launches0 = ctx.launch_count
    gpu = {
        "stereo_matchGrid_points_600": lambda: M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], [], ctx=ctx),
        "stereo_matchGrid_lines_200": lambda: M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], [], ctx=ctx),
        "temporal_match_points_600": lambda: M.match(prev.pdesc_l, curr.pdesc_l, 0.9, [], ctx=ctx),
        "temporal_match_lines_200": lambda: M.match(prev.ldesc_l, curr.ldesc_l, 0.9, [], ctx=ctx),
    }
    ref = oracle.ref if oracle.ref.available() else None
    port = oracle.port
    if ref:
        ref.set_threads(1)
        cpu = {
            "stereo_matchGrid_points_600": lambda: ref.match_grid_points(a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["rows"], a["cols"], a["d2"], a["win"], 0.9, True),
            "stereo_matchGrid_lines_200": lambda: ref.match_grid_lines(b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["rows"], b["cols"], b["d2"], b["dirs2"], 0.75, b["win"], 0.9, True),
            "temporal_match_points_600": lambda: ref.match(prev.pdesc_l, curr.pdesc_l, 0.9, True, True),
            "temporal_match_lines_200": lambda: ref.match(prev.ldesc_l, curr.ldesc_l, 0.9, True, True),
        }
    else:
        cpu = {
            "stereo_matchGrid_points_600": lambda: port.match_grid_points(a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["rows"], a["cols"], a["d2"], a["win"], 0.9, True),
            "stereo_matchGrid_lines_200": lambda: port.match_grid_lines(b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["rows"], b["cols"], b["d2"], b["dirs2"], 0.75, b["win"], 0.9, True),
            "temporal_match_points_600": lambda: port.match(prev.pdesc_l, curr.pdesc_l, 0.9, True),
            "temporal_match_lines_200": lambda: port.match(prev.ldesc_l, curr.ldesc_l, 0.9, True),
        }
    for k in gpu:
        out[k] = {"gpu_ms": _median_ms(gpu[k]), "cpu_ms": _median_ms(cpu[k], reps=10, warm=2)}
    out["per_frame_total_gpu_ms"] = sum(v["gpu_ms"] for v in out.values() if isinstance(v, dict))
    out["per_frame_total_cpu_ms"] = sum(v["cpu_ms"] for v in out.values() if isinstance(v, dict))
    out["cpu_kind"] = "reference (matching.cpp, its own 2 std::async threads for match)" if ref else "port (scalar C)"
    out["note"] = ("wall clock per host-buffer call incl. ctypes, pinned staging, H2D, kernels, D2H and the "
                   "stream sync; includes the Python binding overhead on both sides")
    out["gpu_launches"] = ctx.launch_count - launches0
    return out


def replay_throughput(ctx, n_frames: int) -> dict:
    """Config 3: n_frames stereo frames, stereo matchGrid (points + lines) and temporal match (points +
    lines) per frame, one launch per stage."""
    import torch
    t0 = time.perf_counter()
    rp = synth.make_replay(synth.SEED0 + 3, n_frames)
    gen_s = time.perf_counter() - t0
    gjobs = replay.stereo_grid_jobs(rp)
    tjobs = replay.temporal_match_jobs(rp)
    pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
    arena, coords, cs, ci, dirs = pin(rp.arena), pin(rp.coords), pin(rp.cell_start), pin(rp.cell_items), pin(rp.dirs2)
    m_in = pin(np.full(rp.n_m, -1, np.int32))
    m_out_g = pin(np.empty(rp.n_m, np.int32)); c_out_g = pin(np.empty(len(gjobs), np.int32))
    m_out_t = pin(np.empty(rp.n_m, np.int32)); c_out_t = pin(np.empty(len(tjobs), np.int32))
    gb, tb = replay.MatchBatch(ctx), replay.MatchBatch(ctx)

    def e2e():
        gb.set_match_grid(arena, coords, cs, ci, dirs, 48, 64, gjobs, 0.9, 0.75, True, m_in)
        tb.set_match(arena, tjobs, 0.9, True, m_in)
        gb.run(); tb.run()
        gb.fetch(m_out_g, c_out_g); tb.fetch(m_out_t, c_out_t)

    e2e()  # also warms allocations
    launches0 = ctx.launch_count
    t = time.perf_counter(); e2e(); e2e_s = time.perf_counter() - t
    stream = torch.cuda.ExternalStream(ctx._lib.plm_ctx_stream(ctx.handle)) if False else None  # own stream; timed by wall + sync
    # device-resident: arenas already uploaded by set_*; time run() only
    ctx.synchronize()
    reps = 5
    t = time.perf_counter()
    for _ in range(reps):
        gb.run(); tb.run()
    ctx.synchronize()
    dev_s = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps):
        gb.run()
    ctx.synchronize()
    grid_s = (time.perf_counter() - t) / reps
    launches = ctx.launch_count - launches0
    pairs_t = float(np.sum(tjobs["n1"].astype(np.float64) * tjobs["n2"]))
    matched = int((np.asarray(c_out_g) > 0).sum() // 2)
    return {
        "frames": n_frames, "generation_s": gen_s,
        "device_resident": {"frames_per_s": n_frames / dev_s, "ms_total": dev_s * 1e3,
                            "stereo_grid_ms": grid_s * 1e3, "temporal_match_ms": (dev_s - grid_s) * 1e3,
                            "temporal_unique_pairs_per_s": pairs_t / max(dev_s - grid_s, 1e-9),
                            "temporal_reference_equivalent_pairs_per_s": 2 * pairs_t / max(dev_s - grid_s, 1e-9)},
        "e2e": {"frames_per_s": n_frames / e2e_s, "ms_total": e2e_s * 1e3,
                "h2d_bytes": gb.h2d_bytes + tb.h2d_bytes, "d2h_bytes": gb.d2h_bytes + tb.d2h_bytes,
                "note": "pinned host arenas -> device, 4 stages, match vectors + counts back to pinned host"},
        "frames_with_stereo_matches": matched, "gpu_launches": launches,
        "stage_order": "stereo matchGrid(points, lines) batch, then temporal match(points, lines) batch; stages "
                       "use the full per-frame descriptor sets (no stereo-filter compaction in between)",
    }


def replay_cpu_baseline(n_frames: int, seconds: float = 8.0) -> dict:
    """Frame-parallel thread pool over the reference build on a bounded number of frames."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    cores = len(os.sched_getaffinity(0))
    rp = synth.make_replay(synth.SEED0 + 3, n_frames)
    gjobs = replay.stereo_grid_jobs(rp)
    tjobs = replay.temporal_match_jobs(rp)
    ref_ok = oracle.ref.available()
    if ref_ok:
        oracle.ref.set_threads(1)
    # the reference build keeps Config in a process-wide singleton; set it once, call the raw entry points
    if ref_ok:
        oracle.ref.set_config(best_lr=True, lr_parallel=False, min_ratio_12p=0.9, line_sim_th=0.75)
    eng = oracle.port  # ctypes releases the GIL; the C port is re-entrant (the reference's Config is not)

    def frame(f):
        for jb in gjobs[2 * f:2 * f + 2]:
            n1, n2 = int(jb["n1"]), int(jb["n2"])
            cpq = 4 if jb["is_lines"] else 2
            coords = rp.coords[jb["off_coords"]:jb["off_coords"] + n1 * cpq].reshape(n1, cpq)
            cs = rp.cell_start[jb["off_cell_start"]:jb["off_cell_start"] + 3073]
            ci = rp.cell_items[jb["off_cell_items"]:jb["off_cell_items"] + cs[-1]]
            d1 = rp.arena[jb["off1"]:jb["off1"] + n1]; d2 = rp.arena[jb["off2"]:jb["off2"] + n2]
            if jb["is_lines"]:
                dirs = rp.dirs2[jb["off_dirs2"]:jb["off_dirs2"] + 2 * n2].reshape(n2, 2)
                eng.match_grid_lines(coords, d1, cs, ci, 48, 64, d2, dirs, 0.75, jb["win"], 0.9, True)
            else:
                eng.match_grid_points(coords, d1, cs, ci, 48, 64, d2, jb["win"], 0.9, True)
        for jb in tjobs[2 * f:2 * f + 2]:
            if jb["n2"] < 2:
                continue
            eng.match(rp.arena[jb["off1"]:jb["off1"] + jb["n1"]], rp.arena[jb["off2"]:jb["off2"] + jb["n2"]], 0.9, True)

    t = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(frame, range(n_frames)))
    dt = time.perf_counter() - t
    return {"frames_per_s": n_frames / dt, "frames": n_frames, "cores": cores, "kind": "port",
            "sample": f"{n_frames} frames, frame-parallel thread pool over the C restatement "
                      "(popcount-SWAR distance, scalar), all stages of the replay"}


def run(ctx, args) -> dict:
    out = {"frame_latency": frame_latency(ctx)}
    n = int(os.environ.get("PLM_REPLAY_FRAMES", "10000"))
    out["replay"] = replay_throughput(ctx, n)
    if not args.no_cpu_baseline:
        out["replay"]["cpu_baseline"] = replay_cpu_baseline(min(n, 200))
    return out
