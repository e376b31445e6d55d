"""Multi-GPU partitioning logic on CPU: the shard maths itself, and a world_size-2 gloo run in which
each rank counts its own window shard (oracle in place of the GPU) and rank 0 checks that the
concatenated shard tables equal the single-shard table -- the property the N-GPU path relies on
(no collective on the data path; gloo only carries the check)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_balanced_shards_partition_and_balance():
    from longsom_b200.sharding import balanced_window_shards
    rng = np.random.default_rng(0)
    w = rng.lognormal(0, 1.5, 400)
    for n in (1, 2, 3, 4, 8):
        sh = balanced_window_shards(w, n)
        assert len(sh) == n and sh[0][0] == 0 and sh[-1][1] == len(w)
        assert all(sh[i][1] == sh[i + 1][0] for i in range(n - 1))
        loads = [w[a:b].sum() for a, b in sh]
        assert max(loads) <= w.sum() / n + w.max() + 1e-9
    assert balanced_window_shards([], 4) == [(0, 0)] * 4
    assert balanced_window_shards([1.0], 4)[0] == (0, 1) or sum(b - a for a, b in balanced_window_shards([1.0], 4)) == 1


def test_reads_for_windows_is_a_superset():
    from longsom_b200 import synth
    from longsom_b200.batch import make_windows
    from longsom_b200.pipeline import read_ends
    from longsom_b200.sharding import reads_for_windows
    d = synth.generate(seed=31, contig_lens=[150000, 90000], n_genes=12, n_reads=4000, n_cells=50)
    iv = make_windows(d.contig_lens, 50000)
    wt, ws, we = (np.array([i[k] for i in iv], np.int32) for k in range(3))
    b = d.batch
    ends = read_ends(b)
    for lo, hi in ((0, 2), (2, 4), (1, len(iv))):
        sel = set(reads_for_windows(wt, ws, we, lo, hi, b.tid, b.pos, ends).tolist())
        for r in range(b.n_reads):
            overl = any(b.tid[r] == wt[j] and b.pos[r] < we[j] and ends[r] > ws[j] for j in range(lo, hi))
            assert (not overl) or r in sel


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from longsom_b200 import synth
    from longsom_b200.batch import Windows, make_windows
    from longsom_b200.engine import CountParams
    from longsom_b200.pipeline import read_ends
    from longsom_b200.sharding import balanced_window_shards, reads_for_windows, window_weights
    import oracle
    d = synth.generate(seed=33, contig_lens=[160000, 120000], n_genes=14, n_reads=6000, n_cells=80)
    iv = make_windows(d.contig_lens, 50000)
    wt, ws, we = (np.array([i[k] for i in iv], np.int32) for k in range(3))
    b = d.batch
    ends = read_ends(b)
    weights = window_weights(wt, ws, we, b.tid, b.pos, b.l_qseq)
    lo, hi = balanced_window_shards(weights, world)[rank]
    sel = reads_for_windows(wt, ws, we, lo, hi, b.tid, b.pos, ends)
    prm = CountParams(min_bq=20, min_mq=60)
    mine, _ = oracle.pileup_count(b.select(sel), Windows.from_intervals(iv[lo:hi], d.contig_seqs()), prm)
    # gather shard tables on rank 0 (test plumbing only)
    n = torch.tensor([mine.n_sites], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx, 28), dtype=torch.int64)
    pad[:mine.n_sites, 0] = torch.from_numpy(mine.tid.astype(np.int64))
    pad[:mine.n_sites, 1] = torch.from_numpy(mine.pos.astype(np.int64))
    pad[:mine.n_sites, 2:] = torch.from_numpy(mine.counts.astype(np.int64))
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    ok = True
    if rank == 0:
        cat = np.concatenate([bufs[r][:int(sizes[r].item())].numpy() for r in range(world)])
        full, _ = oracle.pileup_count(b, Windows.from_intervals(iv, d.contig_seqs()), prm)
        ok = (cat.shape[0] == full.n_sites and np.array_equal(cat[:, 0], full.tid) and np.array_equal(cat[:, 1], full.pos)
              and np.array_equal(cat[:, 2:], full.counts.astype(np.int64)))
        open(os.path.join(out_dir, "result"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_concatenate(built, tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "result")).read() == "ok"


def test_window_shards_are_slices_and_reproduce_the_whole(built):
    """pipeline.window_shards (the pipelined single-GPU path): consecutive read ranges + window ranges whose oracle
    results concatenate to the unsharded result; a read spanning a cut is in both shards."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import oracle
    from longsom_b200 import synth
    from longsom_b200.batch import Windows, make_windows
    from longsom_b200.engine import CountParams
    from longsom_b200.pipeline import window_shards
    d = synth.generate(seed=5, contig_lens=[300000, 120000], n_genes=25, n_reads=6000, n_cells=60)
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    p = CountParams(min_bq=20, min_mq=60, min_dp=2, min_cc=1)
    full = oracle.pileup_count(d.batch, w, p)[0]
    shards = window_shards(d.batch, w, 4)
    assert sum(ws.n_windows for _, ws in shards) == w.n_windows
    assert sum(b.n_reads for b, _ in shards) >= d.batch.n_reads * 0.9
    parts = [oracle.pileup_count(b, ws, p)[0] for b, ws in shards]
    assert sum(x.n_sites for x in parts) == full.n_sites
    assert np.array_equal(np.concatenate([x.pos for x in parts]), full.pos)
    assert np.array_equal(np.concatenate([x.counts for x in parts]), full.counts)
    assert window_shards(d.batch, w, 1)[0][0] is d.batch


def test_prewarm_hands_out_background_engines(monkeypatch):
    """pipeline.prewarm / take_engine: engines are built in background threads, handed out once, errors surface."""
    import threading
    import time
    from longsom_b200 import pipeline
    made = []

    class FakeEngine:
        def __init__(self, dev):
            time.sleep(0.05)
            if dev == 7:
                raise RuntimeError("no such device")
            self.dev, self.thread = dev, threading.current_thread().name
            made.append(self)
    monkeypatch.setattr(pipeline, "Engine", FakeEngine)
    pipeline.prewarm([0, 1])
    a, b = pipeline.take_engine(0), pipeline.take_engine(1)
    assert (a.dev, b.dev) == (0, 1) and a.thread != threading.current_thread().name
    c = pipeline.take_engine(0)          # pool empty: built on demand, in the caller's thread
    assert c is not a and c.thread == threading.current_thread().name
    pipeline.prewarm([7])
    with pytest.raises(RuntimeError, match="no such device"):
        pipeline.take_engine(7)
    assert len(made) == 3
