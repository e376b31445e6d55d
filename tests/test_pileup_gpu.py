"""K1 parity: CUDA pileup-count (through the C-ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from longsom_b200 import synth
from longsom_b200.batch import ReadBatch, Windows, make_windows
from longsom_b200.engine import CountParams

pytestmark = pytest.mark.gpu


def _compare(got, want):
    assert got.n_sites == want.n_sites, (got.n_sites, want.n_sites)
    assert np.array_equal(got.tid, want.tid)
    assert np.array_equal(got.pos, want.pos)
    assert np.array_equal(got.ref, want.ref)
    if not np.array_equal(got.counts, want.counts):
        bad = np.nonzero((got.counts != want.counts).any(axis=1))[0]
        i = bad[0]
        raise AssertionError("counts differ at %d sites; first tid=%d pos=%d\n got=%s\nwant=%s" % (
            len(bad), got.tid[i], got.pos[i], got.counts[i], want.counts[i]))


CASES = [
    dict(seed=1, contig_lens=[300000, 16600], chrm=True, n_genes=12, n_reads=20000, n_cells=200),
    dict(seed=2, contig_lens=[120000, 90000, 70000], n_genes=30, n_reads=8000, n_cells=50),
    dict(seed=3, contig_lens=[200000], n_genes=6, n_reads=30000, n_cells=1000, n_hot_genes=2, hot_fraction=0.9),
    dict(seed=4, contig_lens=[100000], n_genes=5, n_reads=300, n_cells=20),
    # indel every ~25 bases: dozens of CIGAR pieces per (read, tile) segment (piece-queue overflow path)
    dict(seed=7, contig_lens=[100000], n_genes=5, n_reads=3000, n_cells=30, p_ins=0.02, p_del=0.02),
    # ~6 kb reads: > 5 (read, tile) segments per read, so the segment builder's first-run size estimate is too
    # small and the exact-size relaunch path runs
    dict(seed=8, contig_lens=[500000], n_genes=8, n_reads=2500, n_cells=40, mean_len=6000.0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("prm", [dict(min_bq=20, min_mq=60), dict(min_bq=30, min_mq=0, min_dp=1, min_cc=1),
                                 dict(min_bq=0, min_mq=255, min_dp=3, min_cc=2),
                                 dict(min_bq=20, min_mq=60, min_ac=2)])
def test_synth_parity(engine, case, prm):
    import oracle
    d = synth.generate(**CASES[case])
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    p = CountParams(**prm)
    got = engine.pileup_count(d.batch, w, p)
    want, nal = oracle.pileup_count(d.batch, w, p, threads=4)
    _compare(got, want)
    assert engine.last_stats["n_aligned"] == nal == d.batch.aligned_bases()
    if prm.get("min_dp", 5) == 5:
        assert got.n_sites > 0


def test_odd_windows(engine):
    """Windows that are not on the 50 kb grid, with gaps (the --bed / --bed_out case)."""
    import oracle
    d = synth.generate(seed=9, contig_lens=[150000, 80000], n_genes=14, n_reads=12000, n_cells=80)
    iv = [(0, 1, 777), (0, 777, 40001), (0, 52000, 52001), (0, 60000, 149999), (1, 5, 33333), (1, 40000, 80000)]
    w = Windows.from_intervals(iv, d.contig_seqs())
    p = CountParams(min_bq=20, min_mq=60, min_dp=2, min_cc=2)
    _compare(engine.pileup_count(d.batch, w, p), oracle.pileup_count(d.batch, w, p)[0])


def test_depth_cap(engine):
    """pileup max_depth: small cap so that the drop rule fires."""
    import oracle
    d = synth.generate(seed=5, contig_lens=[120000], n_genes=4, n_reads=20000, n_cells=300, n_hot_genes=1,
                       hot_fraction=0.8)
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    for cap in (50, 500, 3000):
        p = CountParams(min_bq=20, min_mq=60, max_depth=cap)
        got = engine.pileup_count(d.batch, w, p)
        want = oracle.pileup_count(d.batch, w, p)[0]
        _compare(got, want)
    uncapped = oracle.pileup_count(d.batch, w, CountParams(min_bq=20, min_mq=60, max_depth=0))[0]
    assert not np.array_equal(uncapped.counts[:, 0].sum(), want.counts[:, 0].sum())


def test_empty_and_tiny(engine):
    d = synth.generate(seed=6, contig_lens=[60000], n_genes=4, n_reads=200, n_cells=10)
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    p = CountParams(min_bq=20, min_mq=60)
    empty = d.batch.select(np.zeros(0, np.int64))
    got = engine.pileup_count(empty, w, p)
    assert got.n_sites == 0
    w0 = Windows.from_intervals([], d.contig_seqs())
    assert engine.pileup_count(d.batch, w0, p).n_sites == 0
    # everything filtered by MAPQ
    assert engine.pileup_count(d.batch, w, CountParams(min_bq=20, min_mq=61)).n_sites == 0


def test_rejects_bad_input(engine):
    from longsom_b200._lib import LongSomError
    d = synth.generate(seed=6, contig_lens=[60000], n_genes=4, n_reads=200, n_cells=10)
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    b = d.batch
    bad = ReadBatch(b.tid, b.pos[::-1].copy(), b.flag, b.mapq, b.cell, b.cigar_off, b.cigar, b.base_off, b.l_qseq,
                    b.seq4, b.qual)
    with pytest.raises(LongSomError):
        engine.pileup_count(bad, w, CountParams())


def test_few_cells_long_runs(engine):
    """3 cells x 20k reads on one locus: same-cell runs of thousands of segments per tile
    (the kernel must leave its 12-bit packed-counter variant) and multi-part tiles."""
    import oracle
    d = synth.generate(seed=8, contig_lens=[100000], n_genes=3, n_reads=20000, n_cells=3, n_extra_cells=1,
                       n_hot_genes=1, hot_fraction=0.9)
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    p = CountParams(min_bq=20, min_mq=60, min_dp=5, min_cc=2)
    _compare(engine.pileup_count(d.batch, w, p), oracle.pileup_count(d.batch, w, p, threads=4)[0])


def test_results_independent_of_batch_composition(engine):
    """Size-independent property: counting a window from the full batch or from only the reads
    that overlap it gives identical sites (what multi-GPU sharding relies on)."""
    d = synth.generate(seed=12, contig_lens=[260000], n_genes=16, n_reads=15000, n_cells=120)
    iv = make_windows(d.contig_lens, 50000)
    p = CountParams(min_bq=20, min_mq=60)
    full = engine.pileup_count(d.batch, Windows.from_intervals(iv, d.contig_seqs()), p)
    parts = []
    for t, s, e in iv:
        op = d.batch.cigar & 15
        ln = np.where(np.isin(op, [0, 2, 3, 7, 8]), d.batch.cigar >> 4, 0).astype(np.int64)
        span = np.add.reduceat(ln, d.batch.cigar_off[:-1].astype(np.int64))
        sel = np.nonzero((d.batch.pos < e) & (d.batch.pos + span > s))[0]
        parts.append(engine.pileup_count(d.batch.select(sel), Windows.from_intervals([(t, s, e)], d.contig_seqs()), p))
    assert sum(x.n_sites for x in parts) == full.n_sites
    assert np.array_equal(np.concatenate([x.pos for x in parts]), full.pos)
    assert np.array_equal(np.concatenate([x.counts for x in parts]), full.counts)


def test_pipelined_window_shards_match_single_call(engine):
    """pipeline.count_shards_pipelined: window shards on two CUDA contexts of one GPU (overlapped transfers) give,
    concatenated, exactly the single-call result; the shards are array slices of the batch."""
    from longsom_b200.batch import SiteCounts
    from longsom_b200.pipeline import count_shards_pipelined, window_shards
    d = synth.generate(seed=21, contig_lens=[400000, 150000], n_genes=40, n_reads=40000, n_cells=300)
    w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
    p = CountParams(min_bq=20, min_mq=60)
    full = engine.pileup_count(d.batch, w, p)
    shards = window_shards(d.batch, w, 5)
    assert len(shards) >= 3
    outs = [SiteCounts(np.zeros(full.n_sites, np.int32), np.zeros(full.n_sites, np.int32), np.zeros(full.n_sites, np.uint8),
                       np.zeros((full.n_sites, 26), np.uint32)) for _ in shards]
    for _ in range(2):  # second round: contexts reused
        got = count_shards_pipelined(shards, p, outs, device=0, lanes=2)
        assert sum(got) == full.n_sites
        assert np.array_equal(np.concatenate([o.tid[:c] for o, c in zip(outs, got)]), full.tid)
        assert np.array_equal(np.concatenate([o.pos[:c] for o, c in zip(outs, got)]), full.pos)
        assert np.array_equal(np.concatenate([o.counts[:c] for o, c in zip(outs, got)]), full.counts)
    res = count_shards_pipelined(shards, p, None, device=0, lanes=2)   # outputs allocated per shard after its run
    assert np.array_equal(np.concatenate([r.counts for r in res]), full.counts)
