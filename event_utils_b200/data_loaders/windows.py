"""The window tables of the reference's data loaders and their voxelisation in one launch.

BaseVoxelDataset (reference lib/data_loaders/base_dataset.py) turns a recording into items by a table
of (start, end) event indices -- `compute_k_indices` :354-367, `compute_timeblock_indices` :338-352,
`compute_between_frame_indices` :322-336, chosen by `set_voxel_method` :385-417 -- and voxelises one
window per `__getitem__` (:226-320 -> `get_voxel_grid` :433-455).  With windows of 10 k - 100 k events
that is launch-latency bound on a GPU; here the tables are built vectorised and ALL windows of a table
(or any subset, e.g. a batch) are voxelised by one kernel launch (evk_voxel_windows_f32).
"""
import numpy as np
import torch

from .. import _lib
from ..representations import _events as E


def k_event_indices(num_events, k, sliding_window_w=0):
    """'k_events' (base_dataset.py:393-397, :354-367): item i covers events
    [(k - w) i, (k - w) i + k); int(num_events / (k - w)) items -- the last ones may reach past the
    recording, which the reference only notices when the item is fetched (:429-430)."""
    hop = k - sliding_window_w
    length = max(int(num_events / hop), 0)
    starts = hop * np.arange(length, dtype=np.int64)
    return np.stack((starts, starts + k), 1)


def timeblock_indices(ts, t, sliding_window_t=0.0, length=None):
    """'t_seconds' (base_dataset.py:398-402, :338-352): block i ENDS at the first event not before
    t0 + (t - w) i + t (np.searchsorted, memmap_dataset.py:81-83) and -- as in the reference --
    STARTS where block i-1 ended, also when the windows overlap.  ts: sorted timestamps (numpy)."""
    ts = np.asarray(ts)
    t0, duration = ts[0], ts[-1] - ts[0]
    if length is None:
        length = max(int(duration / (t - sliding_window_t)), 0)
    end_times = (t - sliding_window_t) * np.arange(length) + t0 + t
    ends = np.searchsorted(ts, end_times).astype(np.int64)
    starts = np.concatenate((np.zeros(1, np.int64), ends[:-1])) if length else ends
    return np.stack((starts, ends), 1)


def fixed_frames_indices(ts, num_frames):
    """'fixed_frames' (base_dataset.py:403-407): num_frames blocks of (tk - t0) / num_frames seconds."""
    ts = np.asarray(ts)
    return timeblock_indices(ts, (ts[-1] - ts[0]) / num_frames, 0.0, length=int(num_frames))


def between_frame_indices(ts, frame_ts):
    """'between_frames' (base_dataset.py:322-336): for every frame timestamp the events since the
    previous one; the end index is clamped to num_events - 1."""
    ts = np.asarray(ts)
    ends = np.minimum(np.searchsorted(ts, np.asarray(frame_ts)), len(ts) - 1).astype(np.int64)
    starts = np.concatenate((np.zeros(1, np.int64), ends[:-1])) if len(ends) else ends
    return np.stack((starts, ends), 1)


def check_event_indices(event_indices, num_events):
    """get_event_indices' bounds test (base_dataset.py:428-431) for a whole table."""
    idx = np.asarray(event_indices, dtype=np.int64).reshape(-1, 2)
    bad = np.flatnonzero(~((idx[:, 0] >= 0) & (idx[:, 1] <= num_events)))
    if bad.size:
        raise Exception("WARNING: Event indices {},{} out of bounds 0,{}".format(idx[bad[0], 0], idx[bad[0], 1], num_events))
    return idx


def voxelize_event_windows(xs, ys, ts, ps, event_indices, num_bins, sensor_size=(180, 240),
                           combined_voxel_channels=True, device=None):
    """
    get_voxel_grid (base_dataset.py:433-455) of EVERY window of `event_indices` in one kernel launch.
    @param xs, ys, ts, ps the recording (or the part the windows index): float32 tensors, ps in {-1, +1}
        as `preprocess_events` leaves them (:210-224); CUDA tensors are used in place, host tensors are
        uploaded once
    @param event_indices (n, 2) start / end event indices (any of the tables above, or a batch of rows)
    @param combined_voxel_channels True: (n, num_bins, H, W) = events_to_voxel_torch per window;
        False: (n, 2 num_bins, H, W) = cat(events_to_neg_pos_voxel_torch) per window
    @returns float32 CUDA tensor; an empty window gives a zero grid (the reference raises IndexError on
        ts[-1] of an empty slice)
    """
    L = _lib.lib()
    dev = torch.device(device) if device is not None else (xs.device if xs.is_cuda else E.compute_device())
    idx = check_event_indices(event_indices, len(xs))
    if np.any(idx[:, 1] < idx[:, 0]):
        raise ValueError("window end before its start")
    n = idx.shape[0]
    H, W, B = int(sensor_size[0]), int(sensor_size[1]), int(num_bins)
    grids = 1 if combined_voxel_channels else 2
    with torch.cuda.device(dev):
        x, y, t, p = (a.to(device=dev, dtype=torch.float32).contiguous().reshape(-1) for a in (xs, ys, ts, ps))
        out = torch.empty((n, grids * B, H, W), dtype=torch.float32, device=dev)
        if n == 0:
            return out
        pairs = torch.from_numpy(np.ascontiguousarray(idx.reshape(-1))).to(dev)
        oob = _lib.oob_counter(dev)
        flags = _lib.WINDOW_PAIRS | (0 if combined_voxel_channels else _lib.WINDOW_NEGPOS)
        _lib.check(L.evk_voxel_windows_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), _lib.ptr(pairs), n,
                                           int((idx[:, 1] - idx[:, 0]).sum()), B, H, W, flags, _lib.ptr(out), _lib.ptr(oob),
                                           _lib.stream()))
        E.raise_if_oob(oob, "voxel grid", (B, H, W))
    return out
