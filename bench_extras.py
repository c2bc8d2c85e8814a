"""(Part of bench.py, kept outside the product package because its CPU legs load oracle/.)
Side metrics of bench.py for the frame-scale configs of BASELINE.json (1-3): per-call latency of the
stereo / temporal matchers through the host-buffer C ABI, and replay throughput (matched stereo
frames/s) of the batched path, device-resident and end to end.  The CPU figures next to them come
from the oracle builds (reference sources where available) on this box's host cores."""
from __future__ import annotations

import os
import time

import numpy as np

from pl_inertial_slam_b200 import matching as M
from pl_inertial_slam_b200 import replay, synth


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
    out["per_frame_total_gpu_ms"] = sum(v["gpu_ms"] for v in out.values() if isinstance(v, dict) and "gpu_ms" in v)
    out["per_frame_total_cpu_ms"] = sum(v["cpu_ms"] for v in out.values() if isinstance(v, dict) and "cpu_ms" in v)
    out["cpu_kind"] = "reference (matching.cpp, its own 2 std::async threads for match)" if ref else "port (scalar C)"
    out["note"] = ("wall clock per host-buffer call incl. ctypes, pinned staging, H2D, kernels, D2H and the "
                   "stream sync; includes the Python binding overhead on both sides")
    out["gpu_launches"] = ctx.launch_count - launches0
    # the four calls of the frame as ONE frame session (plm_frame_begin / plm_frame_end -> frame_fused_kernel: one copy
    # in, one launch, results stored straight into the pinned host block): through the Python binding, and at C level
    # by tools/latency_bench (no Python in the loop)
    def session():
        with M.FrameSession(ctx):
            for fn in gpu.values():
                fn()
    l0 = ctx.launch_count
    session()
    out["frame_session"] = {"launches_per_frame": ctx.launch_count - l0, "python_ms": _median_ms(session)}
    try:
        import json as _json
        import subprocess
        tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "latency_bench")
        res = subprocess.run([tool, "300"], capture_output=True, text=True, timeout=120)
        for line in res.stdout.splitlines():
            if line.startswith("JSON "):
                out["frame_session"]["c_level"] = _json.loads(line[5:])
    except Exception as e:  # noqa: BLE001
        out["frame_session"]["c_level"] = {"error": repr(e)[:200]}
    # the same four calls at the reference's C++ signature level (StVO::matchGrid / match with cv::Mat, GridStructure,
    # std::vector<int>&): the GPU drop-in (libstvo_gpu.so: GridStructure -> CSR flattening, staging, copies and sync
    # inside the number) next to the reference's own matching.cpp, timed by the same C++ harness, no Python in the loop
    try:
        dropin = oracle.stvo_gpu
        if ref and dropin.available():
            ref.set_config(True, True, 0.9, 0.75)
            dropin.set_config(True, True, 0.9, 0.75)
            sig = {}
            for tag, eng in (("gpu_dropin_us", dropin), ("cpu_reference_us", ref)):
                reps = 100 if tag.startswith("gpu") else 15
                sig[tag] = {
                    "stereo_matchGrid_points_600": eng.time_match_grid(a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["rows"], a["cols"], a["d2"], a["win"], reps=reps),
                    "stereo_matchGrid_lines_200": eng.time_match_grid(b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["rows"], b["cols"], b["d2"], b["win"], dirs2=b["dirs2"], reps=reps),
                    "temporal_match_points_600": eng.time_match(prev.pdesc_l, curr.pdesc_l, 0.9, reps=reps),
                    "temporal_match_lines_200": eng.time_match(prev.ldesc_l, curr.ldesc_l, 0.9, reps=reps),
                }
                sig[tag]["per_frame_total"] = sum(sig[tag].values())
            # ... and the whole frame as ONE launch through StVO::GpuFrame (same types; GridStructure flattening, staging,
            # copy in, launch and synchronisation inside the number)
            _, _, med = dropin.gpu_frame((a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["d2"], a["win"]),
                                         (b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["d2"], b["dirs2"], b["win"]),
                                         (prev.pdesc_l, curr.pdesc_l), (prev.ldesc_l, curr.ldesc_l), a["rows"], a["cols"], 0.9, reps=200)
            sig["gpu_dropin_us"]["one_launch_frame_StVO_GpuFrame"] = med
            out["stvo_signature_level"] = sig
    except Exception as e:  # noqa: BLE001
        out["stvo_signature_level"] = {"error": repr(e)[:200]}
    return out


def replay_throughput(ctx, n_frames: int, rp=None) -> dict:
    """Config 3: n_frames stereo frames, stereo matchGrid (points + lines) and temporal match (points +
    lines) per frame, one launch per stage."""
    import torch
    t0 = time.perf_counter()
    rp = synth.make_replay(synth.SEED0 + 3, n_frames) if rp is None else rp
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


def replay_pipeline(ctx, n_frames: int, rp=None, chunk: int = 320) -> dict:
    """Config 3 through the device-resident frame pipeline (plm_frames_*): raw keypoints / segments /
    descriptors in, per frame the stereo drivers (grid build, matchGrid, gates, compaction, back-projection)
    and the frame-to-frame match on the compacted descriptors; two launches per feature type for the replay."""
    import torch
    from pl_inertial_slam_b200.frames import FrameConfig, FramePipeline, replay_frame_records
    rp = synth.make_replay(synth.SEED0 + 3, n_frames) if rp is None else rp
    kp, ln, rec = replay_frame_records(rp)
    pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
    arena, kp_t, ln_t = pin(rp.arena), pin(kp), pin(ln)
    cfg = FrameConfig()
    pipe = FramePipeline(ctx)
    pipe.upload(arena, kp_t, ln_t, rec, cfg)
    out = pipe.alloc_outputs(pinned=True)
    pipe.run(); pipe.fetch(out)  # warm (allocations, attribute sets)
    launches0 = ctx.launch_count

    def e2e():
        pipe.process(arena, kp_t, ln_t, rec, cfg, out, chunk_frames=chunk)

    def staged():
        pipe.upload(arena, kp_t, ln_t, rec, cfg)
        pipe.run()
        pipe.fetch(out)

    staged()
    t = time.perf_counter(); staged(); staged_s = time.perf_counter() - t
    ctx.synchronize()
    reps = 5
    t = time.perf_counter()
    for _ in range(reps):
        pipe.run()   # one chunk = the whole replay: four launches
    ctx.synchronize()
    dev_s = (time.perf_counter() - t) / reps
    e2e()
    launches1 = ctx.launch_count
    ts = []
    for _ in range(3):
        t = time.perf_counter(); e2e(); ts.append(time.perf_counter() - t)
    e2e_s = float(np.median(ts))
    per_run = (ctx.launch_count - launches1) // 3
    counts = out["counts"].numpy()
    res = {
        "frames": n_frames,
        "device_resident": {"frames_per_s": n_frames / dev_s, "ms_total": dev_s * 1e3},
        "e2e": {"frames_per_s": n_frames / e2e_s, "ms_total": e2e_s * 1e3, "h2d_bytes": pipe.h2d_bytes,
                "d2h_bytes": pipe.d2h_bytes, "chunk_frames": chunk,
                "unpipelined_frames_per_s": n_frames / staged_s,
                "note": "plm_frames_process: pinned host arenas (descriptors, keypoints, segments) -> device, both "
                        "stages, every output (match vectors, kept lists, disparities, 3-D points / lines, counts) "
                        "back to pinned host; chunks of frames pipelined over copy-in / compute / copy-out streams "
                        "(unpipelined = upload, run, fetch one after the other)"},
        "mean_kept_points": float(counts[:, 1].mean()), "mean_kept_lines": float(counts[:, 3].mean()),
        "mean_f2f_point_matches": float(counts[1:, 4].mean()) if n_frames > 1 else 0.0,
        "frames_with_stereo_matches": int((counts[:, 0] > 0).sum()),
        "gpu_launches": per_run,
        "stages": "stereo_frame_kernel (points), stereo_frame_kernel (lines), f2f_match_kernel (points), "
                  "f2f_match_kernel (lines); temporal matching runs on the stereo-filtered descriptors as in "
                  "stereoFrameHandler.cpp:158-207",
    }
    pipe.close()
    return res


def replay_pipeline_cpu_baseline(n_frames: int) -> dict:
    """The same pipeline on the host cores: frame-parallel thread pool over the C restatement of the stereo
    drivers, then of StVO::match on the compacted descriptors."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from pl_inertial_slam_b200.frames import FrameConfig, replay_frame_records
    cores = len(os.sched_getaffinity(0))
    rp = synth.make_replay(synth.SEED0 + 3, n_frames)
    kp, ln, rec = replay_frame_records(rp)
    cfg = FrameConfig()
    port = oracle.port
    port.lib  # load before the pool starts

    def stereo(f):
        r = rec[f]
        sl = lambda off, n, a: a[int(off):int(off) + int(n)]  # noqa: E731
        dpl = sl(r["desc_pl"], r["n_pl"], rp.arena); dll = sl(r["desc_ll"], r["n_ll"], rp.arena)
        p = port.stereo_points(sl(r["kp_l"], r["n_pl"], kp), dpl, sl(r["kp_r"], r["n_pr"], kp),
                               sl(r["desc_pr"], r["n_pr"], rp.arena), cfg.inv_width, cfg.inv_height, cfg.cam)
        q = port.stereo_lines(sl(r["ln_l"], r["n_ll"], ln), dll, sl(r["ln_r"], r["n_lr"], ln),
                              sl(r["desc_lr"], r["n_lr"], rp.arena), cfg.inv_width, cfg.inv_height, cfg.cam)
        return np.ascontiguousarray(dpl[p["kept_i1"]]), np.ascontiguousarray(dll[q["kept_i1"]])

    def f2f(f):
        for k in (0, 1):
            a, b = comp[f - 1][k], comp[f][k]
            if len(a) >= 2 and len(b) >= 2:
                port.match(a, b, 0.9, True)

    t = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        comp = list(ex.map(stereo, range(n_frames)))
        list(ex.map(f2f, range(1, n_frames)))
    dt = time.perf_counter() - t
    return {"frames_per_s": n_frames / dt, "frames": n_frames, "cores": cores, "kind": "port",
            "sample": f"{n_frames} frames, frame-parallel thread pool over the C restatement of the stereo drivers "
                      "and of StVO::match on the compacted descriptors (scalar SWAR popcount)"}


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


def map_to_frame(ctx, world: int = 1, rank: int = 0, host_call: bool = True) -> dict:
    """Config 4 at N GPUs: a 200 000-point / 50 000-line local map (desc1, with projected cell coordinates),
    row-sharded over the ranks, against one frame (600 / 200 features): matchGrid with a +-3 window, then the
    forced brute-force match() fallback on the same vector (mapHandler.cpp:637-650, :752-765), exchanges included.
    Collective: every rank calls it.  Device-resident timings (CUDA events, max over ranks); at N = 1 also the
    host-buffer call."""
    import torch
    import torch.distributed as dist
    from pl_inertial_slam_b200.database import DeviceOps, GridFrame, ShardedMap, shard_bounds
    from pl_inertial_slam_b200 import grid as G
    sp = synth.make_stereo_pair(synth.SEED0 + 4)
    out = {"n_gpus": world}
    dev = torch.device("cuda", torch.cuda.current_device())
    ops = DeviceOps(dev.index)
    for name, n_map, is_lines in (("points_200k_x_600", 200_000, False), ("lines_50k_x_200", 50_000, True)):
        rng = np.random.default_rng(synth.SEED0 + 40 + int(is_lines))
        if not is_lines:
            d1, xy = synth.make_map_points(synth.SEED0 + 4, n_map, sp)
            d2 = sp.pdesc_l
            c = sp.kp_l.astype(np.float64)
            cs, ci = G.csr_from_points(c[:, 0] * synth.INV_W, c[:, 1] * synth.INV_H)
            dirs = None
            coords = xy
        else:
            d1 = synth.rand_desc(rng, n_map)
            s = np.stack([rng.integers(-2, 66, n_map), rng.integers(-2, 50, n_map)], 1)
            coords = np.concatenate([s, s + rng.integers(-8, 9, (n_map, 2))], 1).astype(np.int32)
            d2 = sp.ldesc_l
            from pl_inertial_slam_b200.drivers import line_grid
            cs, ci, dirs = line_grid(sp.ln_l, synth.INV_W, synth.INV_H)
        frame = GridFrame(torch.from_numpy(d2).to(dev), torch.from_numpy(cs).to(dev), torch.from_numpy(ci).to(dev),
                          G.GRID_ROWS, G.GRID_COLS, torch.from_numpy(dirs).to(dev) if dirs is not None else None)
        lo, hi = shard_bounds(n_map, world, rank)
        smap = ShardedMap(n_map, torch.from_numpy(np.ascontiguousarray(d1[lo:hi])).to(dev),
                          torch.from_numpy(np.ascontiguousarray(coords[lo:hi])).to(dev), ops=ops)
        win = np.array([3, 3, 3, 3], np.int32)

        def timed(fn, reps=20):
            for _ in range(3):
                r = fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(3_000_000)   # ~1.5 ms of GPU spin: the host enqueues ahead, the events see device time only
            e0.record()
            for _ in range(reps):
                r = fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), r

        g_ms, (cnt, m12) = timed(lambda: smap.match_grid(frame, win, 0.9, 0.75, True))
        f_ms, (cnt2, m12b) = timed(lambda: smap.match(frame.d2, 0.9, True, m12_inout=m12))
        pairs = float(n_map) * len(d2)
        import zlib
        res = {"matchGrid_device_ms": g_ms, "matchGrid_matches": int(cnt.item()),
               "matchGrid_crc32": zlib.crc32(m12.cpu().numpy().tobytes()) & 0xFFFFFFFF,
               "match_fallback_device_ms": f_ms, "match_fallback_unique_pairs_per_s": pairs / (f_ms * 1e-3),
               "match_fallback_count": int(cnt2.item()),
               "match_fallback_crc32": zlib.crc32(m12b.cpu().numpy().tobytes()) & 0xFFFFFFFF,
               "exchange": "none" if world == 1 else ("peer-memory kernels" if smap.peer is not None else "nccl")}
        if world == 1 and host_call:
            grid_arg = (cs, ci, G.GRID_ROWS, G.GRID_COLS)
            M.Config.minRatio12P = 0.9
            if dirs is None:
                res["matchGrid_host_call_ms"] = _median_ms(lambda: M.matchGrid(coords, d1, grid_arg, d2, win, np.full(n_map, -1, np.int32), ctx=ctx), reps=8, warm=2)
            else:
                res["matchGrid_host_call_ms"] = _median_ms(lambda: M.matchGrid(coords, d1, grid_arg, d2, dirs, win, np.full(n_map, -1, np.int32), ctx=ctx), reps=8, warm=2)
        if smap.peer is not None:
            smap.peer.check()
        out[name] = res
    return out


def _timed_dev(fn, reps=10):
    """Average device time of fn() (ms) with CUDA events on the context's stream."""
    import torch
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def map_landmarks(ctx, cpu: bool = True) -> dict:
    """The producer of config 4's database rows: MapPoint / MapLine::updateAverageDescDir for the whole
    200 000-point + 50 000-line local map (mean 8 observations per landmark) in one batch -- device-resident
    (CUDA events, arenas in HBM) and through the host-buffer call; the reference's own mapFeatures.cpp on one
    host core beside it (bounded sample)."""
    import torch
    from pl_inertial_slam_b200 import mapfeatures as MF
    n_lm = 250_000
    desc, dirs, start = synth.make_landmark_observations(synth.SEED0 + 21, n_lm, mean_obs=8, long_lists=50, long_len=60)
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    d_desc, d_dirs, d_start = t(desc), t(dirs), t(start)
    med_idx = torch.empty(n_lm, dtype=torch.int32, device=dev)
    med_rows = torch.empty((n_lm, 32), dtype=torch.uint8, device=dev)
    med_dir = torch.empty((n_lm, 3), dtype=torch.float64, device=dev)
    run_dev = lambda: MF.dev_med_desc(ctx, d_desc, d_start, med_idx, med_desc=med_rows, dir_obs=d_dirs, med_dir=med_dir)  # noqa: E731
    torch.cuda.synchronize()
    run_dev(); ctx.synchronize()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        run_dev()
    ctx.synchronize()
    dev_ms = (time.perf_counter() - t0) / reps * 1e3
    host_ms = _median_ms(lambda: MF.med_desc_batch(desc, start, dirs, ctx=ctx), reps=5, warm=1)
    n_obs = int(start[-1])
    alg_bytes = 56.0 * n_obs + 4.0 * (n_lm + 1) + 60.0 * n_lm   # descriptors + directions + offsets in, idx + row + dir out
    out = {"landmarks": n_lm, "observations": n_obs, "device_ms": dev_ms, "host_call_ms": host_ms,
           "landmarks_per_s_device": n_lm / (dev_ms * 1e-3), "algorithmic_bytes": alg_bytes,
           "achieved_gb_s": alg_bytes / (dev_ms * 1e-3) / 1e9,
           "note": "device_ms = wall clock around 10 back-to-back enqueues + one stream sync (two launches each)",
           "gpu_launches": 2}
    if cpu:
        import oracle
        m = 20_000
        sub = slice(0, int(start[m]))
        if oracle.ref.available() and hasattr(oracle.ref.lib, "plref_med_desc"):
            t0 = time.perf_counter(); oracle.ref.med_desc(desc[sub], dirs[sub], start[:m + 1]); dt = time.perf_counter() - t0
            kind = "reference"
        else:
            t0 = time.perf_counter(); oracle.port.med_desc(desc[sub], dirs[sub], start[:m + 1]); dt = time.perf_counter() - t0
            kind = "port"
        out["cpu_baseline"] = {"landmarks_per_s": m / dt, "cores": 1, "kind": kind,
                               "sample": f"{m} landmarks built observation by observation (mapFeatures.cpp), one thread"}
    return out


def bow_scoring(ctx, cpu: bool = True) -> dict:
    """The loop-candidate selector in front of config 5 (mapHandler.cpp:3116-3237): DBoW2 transform of new
    keyframes (800 descriptors each, k = 10 / L = 5 vocabulary with ~10^5 words) and the L1 score of one new
    keyframe against a 20 000-keyframe database, device-resident; the reference's own DBoW2 on one host core
    beside it (bounded sample)."""
    import ctypes as C
    import torch
    from pl_inertial_slam_b200 import _lib as L
    from pl_inertial_slam_b200 import bow as B
    fvoc = synth.make_vocabulary(synth.SEED0 + 42, k=10, L=5, ragged=0.02)
    voc = B.Vocabulary.from_flat(fvoc, ctx=ctx)
    n_kf, per = 2000, 800
    feats = synth.vocabulary_features(synth.SEED0 + 43, fvoc, n_kf * per, flip_p=0.05)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_desc = torch.from_numpy(feats).to(dev)
    d_start = torch.arange(0, (n_kf + 1) * per, per, dtype=torch.int32, device=dev)
    ids = torch.zeros(n_kf * per, dtype=torch.int32, device=dev)
    vals = torch.zeros(n_kf * per, dtype=torch.float64, device=dev)
    lens = torch.zeros(n_kf, dtype=torch.int32, device=dev)
    lib = L.load()
    ptr = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731

    def transform():
        L.check(lib.plm_dev_bow_transform(voc._h, ptr(d_desc), n_kf * per, ptr(d_start), n_kf, per, ptr(ids), ptr(vals),
                                          ptr(lens)), "plm_dev_bow_transform")
    torch.cuda.synchronize()
    transform(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        transform()
    ctx.synchronize()
    tr_ms = (time.perf_counter() - t0) / 5 * 1e3
    # database of 20 000 keyframe vectors: the 2000 transformed keyframes replicated 10 x in HBM (distinct
    # addresses: 16 M entry slots x 12 B = 192 MB, larger than the 126 MB L2)
    n_db, rep = 20_000, 10
    db_ids, db_vals = ids.repeat(rep), vals.repeat(rep)
    db_start = torch.arange(n_db, device=dev, dtype=torch.int64) * per
    db_len = lens.repeat(rep).contiguous()
    q_start = torch.zeros(1, dtype=torch.int64, device=dev)
    q_len = lens[:1].contiguous()
    scores = torch.zeros(n_db, dtype=torch.float64, device=dev)

    def score():
        L.check(lib.plm_dev_bow_score(ctx.handle, ptr(ids), ptr(vals), ptr(q_start), ptr(q_len), 1, per, fvoc.n_words, ptr(db_ids),
                                      ptr(db_vals), ptr(db_start), ptr(db_len), n_db, ptr(scores)), "plm_dev_bow_score")
    torch.cuda.synchronize()
    score(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        score()
    ctx.synchronize()
    sc_ms = (time.perf_counter() - t0) / 10 * 1e3
    entries = int(db_len.sum().item())
    # the same database without the 10 copies of the query keyframe itself (each shares all ~780 words with the query:
    # 780 dependent fp64 additions in the reference's order -- the tail of the launch); what remains shares 5-10 words
    db_len_all = db_len
    db_len = db_len_all.clone()
    db_len[::n_kf] = 0
    score(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        score()
    ctx.synchronize()
    sc2_ms = (time.perf_counter() - t0) / 10 * 1e3
    entries2 = int(db_len.sum().item())
    db_len = db_len_all
    score(); ctx.synchronize()
    out = {"vocabulary": {"k": 10, "L": 5, "nodes": fvoc.n_nodes, "words": fvoc.n_words},
           "transform": {"keyframes": n_kf, "descriptors_per_kf": per, "device_ms": tr_ms,
                         "keyframes_per_s": n_kf / (tr_ms * 1e-3), "descriptors_per_s": n_kf * per / (tr_ms * 1e-3)},
           "score": {"database_keyframes": n_db, "database_entries": entries, "device_ms": sc_ms,
                     "scores_per_s": n_db / (sc_ms * 1e-3), "entries_per_s": entries / (sc_ms * 1e-3),
                     # every database id is read once (4 B); values (8 B) only for the few words the two keyframes share
                     "algorithmic_bytes": 4.0 * entries + 20.0 * n_db,
                     "achieved_gb_s": (4.0 * entries + 20.0 * n_db) / (sc_ms * 1e-3) / 1e9,
                     "database_bytes_resident": 12.0 * n_kf * per * rep,
                     "self_score": float(scores[0].item()),
                     "without_query_copies": {"device_ms": sc2_ms, "database_entries": entries2,
                                              "achieved_gb_s": (4.0 * entries2 + 20.0 * n_db) / (sc2_ms * 1e-3) / 1e9}},
           "note": "device times = wall clock around back-to-back enqueues + one stream sync", "gpu_launches": 2}
    if cpu:
        import oracle
        if oracle.ref_dbow.available():
            h = oracle.ref_dbow.from_flat(fvoc)
            m = 20
            t0 = time.perf_counter()
            bows = [oracle.ref_dbow.transform(h, feats[i * per:(i + 1) * per]) for i in range(m)]
            t_tr = (time.perf_counter() - t0) / m
            t0 = time.perf_counter()
            for j in range(2000):
                oracle.ref_dbow.score(h, bows[0], bows[j % m])
            t_sc = (time.perf_counter() - t0) / 2000
            oracle.ref_dbow.destroy(h)
            out["cpu_baseline"] = {"transform_keyframes_per_s": 1.0 / t_tr, "scores_per_s": 1.0 / t_sc, "cores": 1,
                                   "kind": "reference",
                                   "sample": f"{m} keyframes transformed, 2000 scores; the reference's DBoW2 through "
                                             "ctypes (score includes rebuilding two std::map BowVectors per call)"}
    voc.close()
    return out


def loop_closure_per_pair(ctx, cpu: bool = True) -> dict:
    """Config 5 in the reference's own form (SURVEY 8d "Mode A"): isLoopClosure runs StVO::match per keyframe PAIR, each
    with both ratio tests and its own mutual check (mapHandler.cpp:3325-3378).  One query keyframe (800 descriptors)
    against every keyframe of a resident 2000-keyframe database in one batched launch; the reference's match() on one
    core (its own two std::async threads) beside it."""
    import torch
    from pl_inertial_slam_b200.database import KeyframeDB
    n_kf, per = int(os.environ.get("PLM_MODE_A_KFS", "2000")), 800       # 20000 = the full config-5 database
    rng = np.random.default_rng(synth.SEED0 + 55)
    rows = synth.rand_desc(rng, n_kf * per)
    kf_start = np.arange(n_kf + 1, dtype=np.int64) * per
    db = KeyframeDB(rows, kf_start, device=torch.cuda.current_device(), q_cap=per, ctx=ctx)
    q = synth.flip_bits(rng, rows[77 * per:78 * per], 0.05)      # revisits keyframe 77
    counts, _ = db.match_all(q, 0.9, True)
    assert int(np.argmax(counts)) == 77
    launches0 = ctx.launch_count
    reps = 5
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        db.batch.run()
    ctx.synchronize()
    dev_ms = (time.perf_counter() - t0) / reps * 1e3
    call_ms = _median_ms(lambda: db.match_all(q, 0.9, True), reps=5, warm=1)
    pairs = float(n_kf) * per * per
    out = {"keyframes": n_kf, "descriptors_per_kf": per, "device_ms": dev_ms, "host_call_ms": call_ms,
           "keyframe_pairs_per_s": n_kf / (dev_ms * 1e-3), "unique_pairs_per_s": pairs / (dev_ms * 1e-3),
           "reference_equivalent_pairs_per_s": 2 * pairs / (dev_ms * 1e-3), "best_match_count": int(counts.max()),
           "gpu_launches": (ctx.launch_count - launches0) // (reps + 6),
           "note": "device_ms = kernels of one query against all keyframes (both directions + mutual check per pair), "
                   "database resident; host_call_ms adds the query upload and the read-back of the 2000 counts"}
    if cpu:
        import oracle
        eng = oracle.ref if oracle.ref.available() else oracle.port
        m = 20
        t0 = time.perf_counter()
        for j in range(m):
            eng.match(q, rows[j * per:(j + 1) * per], np.float32(0.9), True)
        dt = (time.perf_counter() - t0) / m
        out["cpu_baseline"] = {"keyframe_pairs_per_s": 1.0 / dt, "reference_equivalent_pairs_per_s": 2.0 * per * per / dt,
                               "cores": 2, "kind": "reference" if oracle.ref.available() else "port",
                               "sample": f"{m} keyframe pairs, StVO::match with its two std::async threads"}
    return out


def run(ctx, args) -> dict:
    out = {"frame_latency": frame_latency(ctx)}
    for name, fn in (("loop_closure_per_pair", loop_closure_per_pair), ("map_landmarks", map_landmarks),
                     ("bow_scoring", bow_scoring)):
        try:
            out[name] = fn(ctx, cpu=not args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)}
    n = int(os.environ.get("PLM_REPLAY_FRAMES", "10000"))
    rp = synth.make_replay(synth.SEED0 + 3, n)
    out["replay"] = replay_pipeline(ctx, n, rp)
    out["replay_stage_batches"] = replay_throughput(ctx, n, rp)
    if not args.no_cpu_baseline:
        out["replay"]["cpu_baseline"] = replay_pipeline_cpu_baseline(min(n, 200))
        out["replay_stage_batches"]["cpu_baseline"] = replay_cpu_baseline(min(n, 200))
    return out
