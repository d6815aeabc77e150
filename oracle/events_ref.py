"""ORACLE -- test infrastructure only; never imported by the product path.

Vectorised numpy restatement of the reference's event -> frame code:
  mvsecRectifyEvents             /root/reference/datasets/MVSEC/utils.py:31-56
  mvsecCumulateSpikesIntoFrames  /root/reference/datasets/MVSEC/utils.py:215-281
Pinned by tests/test_oracle_events.py, which executes the reference's own functions (per-event Python loops) from
/root/reference on small synthetic streams and requires identical frames.
Deviation (documented, also in the CUDA kernel): rectified coordinates equal to 346 / 260 pass the reference's field-of-view
filter (`<=`, utils.py:52-55) and would then raise IndexError in its accumulation loop; here such events are dropped.
"""
import numpy as np

FRAME_W, FRAME_H, LIDAR_FPS = 346, 260, 20


def rectify_events(events, x_map, y_map):
    """utils.py:31-56.  events float64 [n,4] = (x, y, t, polarity)."""
    x = events[:, 0].astype(np.int64)
    y = events[:, 1].astype(np.int64)
    out = np.stack([x_map[y, x], y_map[y, x], events[:, 2], events[:, 3]], axis=1)
    keep = (out[:, 0] >= 0) & (out[:, 0] <= FRAME_W) & (out[:, 1] >= 0) & (out[:, 1] <= FRAME_H)
    return out[keep]


def frame_boundaries(n_chunks, num_frames_per_depth_map):
    """The (start, end) timestamps of utils.py:256-258, evaluated with the same float64 expression order."""
    fps = num_frames_per_depth_map * LIDAR_FPS
    starts, ends = [], []
    for numchunk in range(n_chunks):
        for numframe in range(num_frames_per_depth_map):
            starts.append(numchunk * num_frames_per_depth_map * 1 / fps + numframe * 1 / fps)
            ends.append(numchunk * num_frames_per_depth_map * 1 / fps + numframe * 1 / fps + 1 / fps)
    return np.array(starts, dtype=np.float64), np.array(ends, dtype=np.float64)


def cumulate_spikes_into_frames(events, n_chunks, num_frames_per_depth_map=1):
    """utils.py:215-281 without the depth-map bookkeeping.  Returns float64 [n_chunks, nfpdm, 2, 260, 346]."""
    ev = np.array(events, dtype=np.float64, copy=True)
    ev[:, 2] -= ev[0, 2]                                   # utils.py:251-252
    starts, ends = frame_boundaries(n_chunks, num_frames_per_depth_map)
    frames = np.zeros((len(starts), 2, FRAME_H, FRAME_W), dtype=np.float64)
    for f, (s, e) in enumerate(zip(starts, ends)):
        sel = ev[(ev[:, 2] > s) & (ev[:, 2] < e)]          # both strict (utils.py:259)
        xs = sel[:, 0].astype(np.int64)                    # int() truncation (utils.py:262-263)
        ys = sel[:, 1].astype(np.int64)
        ok = (xs >= 0) & (xs < FRAME_W) & (ys >= 0) & (ys < FRAME_H)
        ch = np.where(sel[:, 3] == 1, 0, 1)                # utils.py:266-269
        np.add.at(frames[f], (ch[ok], ys[ok], xs[ok]), 1.0)
    return frames.reshape(n_chunks, num_frames_per_depth_map, 2, FRAME_H, FRAME_W)


def synthetic_events(n, n_chunks, nfpdm, seed=0, raw=False):
    """Sorted synthetic stream with events exactly on frame boundaries, outside the time range and at the image border."""
    rng = np.random.default_rng(seed)
    fps = nfpdm * LIDAR_FPS
    t = np.sort(rng.uniform(-0.01, n_chunks * nfpdm / fps + 0.01, n))
    t[0] = 0.0
    k = rng.integers(1, n_chunks * nfpdm, 8)
    t[rng.integers(1, n, 8)] = k / fps                     # exactly on a boundary: counted by neither neighbour
    t = np.sort(t)
    if raw:
        x = rng.integers(0, FRAME_W, n).astype(np.float64)
        y = rng.integers(0, FRAME_H, n).astype(np.float64)
    else:
        x = rng.uniform(0, FRAME_W - 1e-9, n)
        y = rng.uniform(0, FRAME_H - 1e-9, n)
    p = rng.choice([1.0, -1.0, 0.0], n, p=[0.5, 0.45, 0.05])
    return np.stack([x, y, t + 123.456, p], axis=1)
