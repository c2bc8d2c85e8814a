"""Host-side layout helpers (no GPU): the frame-by-frame arenas the replay hands to plm_frames_process."""
import numpy as np

from pl_inertial_slam_b200 import synth
from pl_inertial_slam_b200.frames import replay_frame_records


def test_replay_frame_records_are_frame_major():
    """Keypoints / segments of frame f are [left | right] right after frame f - 1's, so a chunk of consecutive frames is
    one contiguous slice of each arena; the records point at exactly the rows the generator produced."""
    rp = synth.make_replay(synth.SEED0 + 11, 9)
    kp, ln, rec = replay_frame_records(rp)
    pb = np.concatenate([[0], np.cumsum(rp.n_pts)])
    lb = np.concatenate([[0], np.cumsum(rp.n_lines)])
    kp_l, kp_r = np.asarray(rp.kp_l, np.float32).reshape(-1, 2), np.asarray(rp.kp_r, np.float32).reshape(-1, 2)
    ln_l, ln_r = np.asarray(rp.ln_l, np.float32).reshape(-1, 4), np.asarray(rp.ln_r, np.float32).reshape(-1, 4)
    assert kp.shape == (2 * pb[-1], 2) and ln.shape == (2 * lb[-1], 4)
    for f in range(rp.n_frames):
        n, m = int(rp.n_pts[f]), int(rp.n_lines[f])
        a, b = int(rec["kp_l"][f]), int(rec["kp_r"][f])
        assert a == 2 * pb[f] and b == a + n                      # [left | right] of the frame, frames back to back
        assert np.array_equal(kp[a:a + n], kp_l[pb[f]:pb[f] + n]) and np.array_equal(kp[b:b + n], kp_r[pb[f]:pb[f] + n])
        c, d = int(rec["ln_l"][f]), int(rec["ln_r"][f])
        assert c == 2 * lb[f] and d == c + m
        assert np.array_equal(ln[c:c + m], ln_l[lb[f]:lb[f] + m]) and np.array_equal(ln[d:d + m], ln_r[lb[f]:lb[f] + m])
        assert rec["n_pl"][f] == rec["n_pr"][f] == n and rec["n_ll"][f] == rec["n_lr"][f] == m
