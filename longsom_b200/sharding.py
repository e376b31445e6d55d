"""Multi-GPU partitioning by coverage-balanced genomic bins.

The reference parallelises over 50 kb windows with no exchange between workers
(BaseCellCounter.py:392-402); the B200 equivalent gives every GPU a contiguous run of
windows whose aligned-base weight is ~1/N of the total.  A read that overlaps a shard
boundary is handed to both neighbours (each shard only emits sites inside its own windows),
so no collective is needed and per-shard outputs concatenate in window order.
"""
import numpy as np


def balanced_window_shards(weights, n_shards):
    """Cut windows [0, W) into n_shards contiguous ranges with ~equal total weight.

    Returns a list of (w_lo, w_hi).  Empty ranges are possible when n_shards > W."""
    w = np.asarray(weights, dtype=np.float64)
    W = w.shape[0]
    if n_shards <= 1 or W == 0:
        return [(0, W)] + [(W, W)] * (max(n_shards, 1) - 1)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for s in range(1, n_shards):
        target = total * s / n_shards
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, cuts[-1]), W)
        # pick the neighbour closer to the target
        if c > cuts[-1] and c <= W and abs(cum[c - 1] - target) < abs(cum[min(c, W)] - target):
            c -= 1
        cuts.append(max(c, cuts[-1]))
    cuts.append(W)
    return [(cuts[i], cuts[i + 1]) for i in range(n_shards)]


def window_weights(win_tid, win_start, win_end, read_tid, read_pos, read_weight):
    """Aligned-base weight per window from a cheap per-read pre-pass: each read's weight is
    credited to the window that contains its start (reads before the first window of their
    contig are credited to the next one)."""
    win_tid = np.asarray(win_tid, np.int64)
    key_w = (win_tid << 32) | np.asarray(win_end, np.int64)
    key_r = (np.asarray(read_tid, np.int64) << 32) | np.asarray(read_pos, np.int64)
    idx = np.searchsorted(key_w, key_r, side="right")
    idx = np.minimum(idx, len(key_w) - 1)
    out = np.zeros(len(key_w), np.float64)
    np.add.at(out, idx, np.asarray(read_weight, np.float64))
    return out


def reads_for_windows(win_tid, win_start, win_end, w_lo, w_hi, read_tid, read_pos, read_end):
    """Indices (ascending) of the reads that can overlap windows [w_lo, w_hi): pos < last end and
    end > first start, contig-aware.  read_end may be an upper bound (extra reads are harmless)."""
    if w_hi <= w_lo:
        return np.zeros(0, np.int64)
    lo_key = (int(win_tid[w_lo]) << 32) | int(win_start[w_lo])
    hi_key = (int(win_tid[w_hi - 1]) << 32) | int(win_end[w_hi - 1])
    rt = np.asarray(read_tid, np.int64)
    start_key = (rt << 32) | np.asarray(read_pos, np.int64)
    end_key = (rt << 32) | np.asarray(read_end, np.int64)
    return np.nonzero((start_key < hi_key) & (end_key > lo_key))[0].astype(np.int64)
