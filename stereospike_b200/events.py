"""Event stream -> network input on the device, replacing the per-event Python loops of the reference's data layer
(``mvsecRectifyEvents`` / ``mvsecCumulateSpikesIntoFrames``, datasets/MVSEC/utils.py:31-56,215-281).
The frames are produced directly as the packed u8 ``[T, B, H, W, 4]`` tensor the first fused block reads (left camera in
channels 0-1, right camera in 2-3), ready for ``model.forward_seq``; the dense float frames never exist.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream

FRAME_W, FRAME_H, LIDAR_FPS = 346, 260, 20


def frame_boundaries(n_chunks, num_frames_per_depth_map):
    """(start, end) timestamps of every frame, evaluated exactly like utils.py:256-258 (float64, same expression order)."""
    fps = num_frames_per_depth_map * LIDAR_FPS
    starts, ends = [], []
    for numchunk in range(n_chunks):
        for numframe in range(num_frames_per_depth_map):
            starts.append(numchunk * num_frames_per_depth_map * 1 / fps + numframe * 1 / fps)
            ends.append(numchunk * num_frames_per_depth_map * 1 / fps + numframe * 1 / fps + 1 / fps)
    return np.array(starts, dtype=np.float64), np.array(ends, dtype=np.float64)


def cumulate_spikes_into_frames(events_left, events_right=None, n_chunks=1, num_frames_per_depth_map=1, maps_left=None,
                                maps_right=None, status=None):
    """events_*: CUDA float64 ``[n, 4]`` = (x, y, t, polarity), time-sorted, as the reference stores them (each stream's clock
    is shifted by its own first timestamp, utils.py:251-252).  maps_*: optional (x_map, y_map) float64 ``[260, 346]``
    rectification tables.  Returns u8 ``[T = num_frames_per_depth_map, B = n_chunks, 260, 346, 4]``."""
    dev = events_left.device
    if not events_left.is_cuda:
        raise RuntimeError('stereospike_b200: events must be CUDA tensors -- the hot path has no CPU fallback')
    F = n_chunks * num_frames_per_depth_map
    starts, ends = frame_boundaries(n_chunks, num_frames_per_depth_map)
    starts, ends = torch.from_numpy(starts).to(dev), torch.from_numpy(ends).to(dev)
    counts = torch.zeros((F, FRAME_H, FRAME_W, 4), dtype=torch.int32, device=dev)
    L = _lib.lib()
    for cam, (ev, maps) in enumerate(((events_left, maps_left), (events_right, maps_right))):
        if ev is None or ev.numel() == 0:
            continue
        assert ev.dtype == torch.float64 and ev.dim() == 2 and ev.shape[1] == 4
        ev = ev.contiguous()
        xm = maps[0].contiguous() if maps is not None else None
        ym = maps[1].contiguous() if maps is not None else None
        if maps is None:
            t0 = float(ev[0, 2])
        else:
            # the reference rectifies first and then shifts the clock by the first SURVIVING event (utils.py:52-55,251-252)
            xi, yi = ev[:, 0].long(), ev[:, 1].long()
            xr, yr = xm[yi, xi], ym[yi, xi]
            keep = (xr >= 0) & (xr <= FRAME_W) & (yr >= 0) & (yr <= FRAME_H)
            idx = torch.nonzero(keep)
            if idx.numel() == 0:
                continue
            t0 = float(ev[int(idx[0]), 2])
        rc = L.ss_events_accumulate(_ptr(ev), ev.shape[0], _ptr(xm), _ptr(ym), ctypes.c_double(t0), _ptr(starts), _ptr(ends), F,
                                    FRAME_H, FRAME_W, cam, _ptr(counts), _stream())
        _lib.check(rc, 'ss_events_accumulate')
    out = torch.empty((num_frames_per_depth_map, n_chunks, FRAME_H, FRAME_W, 4), dtype=torch.uint8, device=dev)
    _lib.check(L.ss_events_pack(_ptr(counts), n_chunks, num_frames_per_depth_map, FRAME_H, FRAME_W, _ptr(out), _ptr(status),
                                _stream()), 'ss_events_pack')
    return out
