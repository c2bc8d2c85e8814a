"""Batched replay (config 3): every job of the stereo matchGrid batch and of the temporal match batch
must equal the oracle run on that job alone."""
import numpy as np
import pytest

import oracle
from pl_inertial_slam_b200 import synth

pytestmark = pytest.mark.gpu
port = oracle.port


def _grid_job_oracle(rp, jb, ratio, best_lr, m_in):
    n1, n2 = int(jb["n1"]), int(jb["n2"])
    cpq = 4 if jb["is_lines"] else 2
    coords = rp.coords[jb["off_coords"]:jb["off_coords"] + n1 * cpq].reshape(n1, cpq)
    cs = rp.cell_start[jb["off_cell_start"]:jb["off_cell_start"] + 3073]
    ci = rp.cell_items[jb["off_cell_items"]:jb["off_cell_items"] + cs[-1]]
    d1 = rp.arena[jb["off1"]:jb["off1"] + n1]
    d2 = rp.arena[jb["off2"]:jb["off2"] + n2]
    if jb["is_lines"]:
        dirs = rp.dirs2[jb["off_dirs2"]:jb["off_dirs2"] + 2 * n2].reshape(n2, 2)
        return port.match_grid_lines(coords, d1, cs, ci, 48, 64, d2, dirs, 0.75, jb["win"], ratio, best_lr, m_in)
    return port.match_grid_points(coords, d1, cs, ci, 48, 64, d2, jb["win"], ratio, best_lr, m_in)


@pytest.mark.parametrize("best_lr", [True, False])
def test_replay_batches(plm_lib, best_lr):
    from pl_inertial_slam_b200 import replay
    rp = synth.make_replay(synth.SEED0 + 3, 24, mean_pts=300, sd_pts=40, mean_lines=100, sd_lines=20)
    m_in = np.full(rp.n_m, -1, np.int32)

    gb = replay.MatchBatch()
    gjobs = replay.stereo_grid_jobs(rp)
    gb.set_match_grid(rp.arena, rp.coords, rp.cell_start, rp.cell_items, rp.dirs2, 48, 64, gjobs, 0.9, 0.75, best_lr, m_in)
    for _ in range(2):  # run() restores the IN vector, so a second run gives the same answer
        gb.run()
        m12, counts = gb.fetch()
    for j, jb in enumerate(gjobs):
        n_o, m_o = _grid_job_oracle(rp, jb, 0.9, best_lr, None)
        got = m12[jb["off_m"]:jb["off_m"] + jb["n1"]]
        assert counts[j] == n_o and (got == m_o).all(), j
    assert gb.h2d_bytes > rp.arena.nbytes and gb.d2h_bytes == 4 * (rp.n_m + len(gjobs))

    tb = replay.MatchBatch()
    tjobs = replay.temporal_match_jobs(rp)
    tb.set_match(rp.arena, tjobs, 0.9, best_lr, m_in)
    tb.run()
    m12, counts = tb.fetch()
    assert counts[0] == np.iinfo(np.int32).min and counts[1] == np.iinfo(np.int32).min  # frame 0: no predecessor
    for j, jb in enumerate(tjobs):
        if jb["n2"] < 2:
            continue
        d1 = rp.arena[jb["off1"]:jb["off1"] + jb["n1"]]
        d2 = rp.arena[jb["off2"]:jb["off2"] + jb["n2"]]
        n_o, m_o = port.match(d1, d2, 0.9, best_lr)
        got = m12[jb["off_m"]:jb["off_m"] + jb["n1"]]
        assert counts[j] == n_o and (got == m_o).all(), j


def test_batch_match_stale_inout_and_small_batch(plm_lib):
    """A 3-job batch (slices > 1 per job) with stale IN entries, like the map fallback call sites."""
    from pl_inertial_slam_b200 import _lib as L
    from pl_inertial_slam_b200 import replay
    rng = np.random.default_rng(12)
    sizes = [(700, 650), (40, 1500), (2, 2)]
    rows, jobs, m_in, off, off_m = [], np.zeros(3, L.PAIR_JOB_DTYPE), [], 0, 0
    for j, (n1, n2) in enumerate(sizes):
        d2 = synth.tie_stress_desc(rng, n2) if j == 1 else synth.rand_desc(rng, n2)
        d1 = synth.tie_stress_desc(rng, n1) if j == 1 else synth.rand_desc(rng, n1)
        rows += [d1, d2]
        jobs[j] = (off, off + n1, off_m, n1, n2)
        st = np.full(n1, -1, np.int32)
        st[::3] = rng.integers(0, n2, len(st[::3]))
        m_in.append(st)
        off += n1 + n2
        off_m += n1
    arena = np.concatenate(rows)
    m_in = np.concatenate(m_in)
    b = replay.MatchBatch()
    b.set_match(arena, jobs, 0.75, True, m_in)
    b.run()
    m12, counts = b.fetch()
    for j, jb in enumerate(jobs):
        d1 = arena[jb["off1"]:jb["off1"] + jb["n1"]]
        d2 = arena[jb["off2"]:jb["off2"] + jb["n2"]]
        n_o, m_o = port.match(d1, d2, 0.75, True, m12=m_in[jb["off_m"]:jb["off_m"] + jb["n1"]])
        assert counts[j] == n_o and (m12[jb["off_m"]:jb["off_m"] + jb["n1"]] == m_o).all(), j
