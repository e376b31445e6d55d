"""Parity on the shapes of BASELINE.json's configs (scaled so the CPU oracle finishes in seconds)
and size-independent properties on a larger batch."""
import numpy as np
import pytest

from longsom_b200 import synth
from longsom_b200.batch import ReadBatch, SiteCounts, Windows, make_windows
from longsom_b200.engine import CountParams

pytestmark = pytest.mark.gpu


def _windows(d):
    return Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())


def _same(a, b):
    assert a.n_sites == b.n_sites
    assert np.array_equal(a.tid, b.tid) and np.array_equal(a.pos, b.pos) and np.array_equal(a.ref, b.ref)
    assert np.array_equal(a.counts, b.counts)


@pytest.mark.parametrize("name,scale", [("C1", 0.1), ("C2", 0.01), ("C4", 0.03)])
def test_config_shapes_match_oracle(engine, name, scale):
    import oracle
    d = synth.generate(**synth.config(name, scale=scale))
    w = _windows(d)
    p = CountParams(min_bq=20, min_mq=60)
    got = engine.pileup_count(d.batch, w, p)
    st = engine.last_stats
    want, nal = oracle.pileup_count(d.batch, w, p, threads=8)
    _same(got, want)
    assert st["n_aligned"] == nal and got.n_sites > 0
    if name == "C4":  # hotspot genes: deep tiles are split into parts merged in HBM
        assert st["n_segments"] / max(1, st["n_tiles"]) > 100


def _compare_in_window_chunks(engine, d, prm, n_chunks, threads=None):
    """GPU table of the whole batch vs the oracle run chunk of windows by chunk of windows (bounds the oracle's
    output buffers at full scale); the GPU table is ordered by (window, position), so the chunks are slices of it."""
    import os
    import oracle
    iv = make_windows(d.contig_lens, 50000)
    seqs = d.contig_seqs()
    got = engine.pileup_count(d.batch, Windows.from_intervals(iv, seqs), prm)
    st = dict(engine.last_stats)
    threads = threads or min(32, os.cpu_count() or 8)
    off, step = 0, (len(iv) + n_chunks - 1) // n_chunks
    for c0 in range(0, len(iv), step):
        want, _ = oracle.pileup_count(d.batch, Windows.from_intervals(iv[c0:c0 + step], seqs), prm, threads=threads)
        n = want.n_sites
        assert off + n <= got.n_sites, "GPU table is short at window %d" % c0
        assert np.array_equal(got.tid[off:off + n], want.tid) and np.array_equal(got.pos[off:off + n], want.pos), c0
        assert np.array_equal(got.ref[off:off + n], want.ref), c0
        assert np.array_equal(got.counts[off:off + n], want.counts), c0
        off += n
    assert off == got.n_sites
    return got, st


def test_c1_full_scale(engine):
    """BASELINE.json configs[0] at scale 1.0: chr21 + chrM, 2e5 reads, 1 000 cells."""
    d = synth.generate(**synth.config("C1", scale=1.0))
    got, st = _compare_in_window_chunks(engine, d, CountParams(min_bq=20, min_mq=60), 4)
    assert got.n_sites > 100000 and st["n_aligned"] == d.batch.aligned_bases()


def test_c2_full_scale(engine):
    """BASELINE.json configs[1] at scale 1.0 (the bench workload): 5e6 reads x ~1.5 kb, 5 000 cells, every one of
    the ~1.7e7 emitted sites bit-compared with the oracle."""
    d = synth.generate(**synth.config("C2", scale=1.0))
    got, st = _compare_in_window_chunks(engine, d, CountParams(min_bq=20, min_mq=60, min_dp=5, min_cc=5), 16)
    assert got.n_sites > 10_000_000 and st["n_events"] > 4_000_000_000


def test_c4_full_scale_real_depth_cap(engine):
    """BASELINE.json configs[3] at scale 1.0 (2.6e6 reads, 10 000 cells) with eight hot genes instead of twenty, so
    that one locus holds > 2.5e5 reads and the REAL pileup max_depth = 200000 drops records."""
    cfg = synth.config("C4", scale=1.0)
    cfg["n_hot_genes"] = 8
    d = synth.generate(**cfg)
    prm = CountParams(min_bq=20, min_mq=60, max_depth=200000)
    got, st = _compare_in_window_chunks(engine, d, prm, 8)
    assert st["n_segments"] / max(1, st["n_tiles"]) > 100  # deep tiles: many parts merged in HBM
    w = _windows(d)
    uncapped = engine.pileup_count(d.batch, w, CountParams(min_bq=20, min_mq=60, max_depth=0))
    assert int(uncapped.counts[:, 0].astype(np.int64).sum()) > int(got.counts[:, 0].astype(np.int64).sum())  # the cap fired


def test_c4_depth_cap_fires_at_a_hotspot(engine):
    """One locus deeper than max_depth (dedicated cap test), cap lowered to keep it small."""
    import oracle
    d = synth.generate(seed=44, contig_lens=[100000], n_genes=3, n_reads=40000, n_cells=2000, n_hot_genes=1,
                       hot_fraction=0.95)
    w = _windows(d)
    p = CountParams(min_bq=20, min_mq=60, max_depth=8000)
    got = engine.pileup_count(d.batch, w, p)
    want, _ = oracle.pileup_count(d.batch, w, p, threads=8)
    _same(got, want)
    uncapped, _ = oracle.pileup_count(d.batch, w, CountParams(min_bq=20, min_mq=60, max_depth=0), threads=8)
    assert uncapped.counts[:, 0].sum() > want.counts[:, 0].sum()  # the cap really dropped records


@pytest.mark.parametrize("n_sites,n_cells,scale", [(1000, 1000, 0.02), (10000, 1000, 0.02), (3000, 5000, 0.02),
                                                   (50000, 5000, 0.2)])
def test_c5_genotyping_sweep(engine, n_sites, n_cells, scale):
    import oracle
    from scipy.stats import betabinom
    cfg = synth.config("C5", scale=scale)
    cfg["n_cells"] = n_cells
    d = synth.generate(**cfg)
    w = _windows(d)
    sc = engine.pileup_count(d.batch, w, CountParams(min_mq=60))
    rng = np.random.default_rng(5)
    idx = np.sort(rng.choice(sc.n_sites, size=min(n_sites, sc.n_sites), replace=False))
    alt = rng.integers(0, 4, size=idx.shape[0]).astype(np.uint8)
    dp, al = engine.genotype_count(sc.tid[idx], sc.pos[idx], alt, n_cells, min_bq=30, min_mq=60)
    odp, oal = oracle.genotype_count(d.batch, sc.tid[idx], sc.pos[idx], alt, n_cells, min_bq=30, min_mq=60)
    assert np.array_equal(dp, odp) and np.array_equal(al, oal)
    ri, ci = np.nonzero(al > 0)
    a2, b2 = 0.2474528917555431, 162.03696139428595
    p = engine.betabinom_sf(al[ri, ci], dp[ri, ci], a2, b2)
    ref = betabinom.sf(al[ri, ci] - 0.001, dp[ri, ci], a2, b2)
    assert np.all(np.abs(p - ref) <= 1e-9 * ref + 1e-12)
    assert np.array_equal(np.round(p, 4), np.round(ref, 4))
    assert len(ri) > 0


def _check_sparse_against_dense(engine, tid, pos, alt, n_cells, dp, al, a2, b2, skip=None):
    """The sparse tuples must be exactly the non-zero entries of the dense tensors, in (site, cell) order, with the
    device-side beta-binomial tails equal to the separately evaluated ones."""
    from scipy.stats import betabinom
    site, cell, sdp, sal, p = engine.genotype_sparse(tid, pos, alt, n_cells, a2, b2, skip_p=skip, min_bq=30, min_mq=60)
    ri, ci = np.nonzero(dp > 0)
    assert np.array_equal(site, ri) and np.array_equal(cell, ci)
    assert np.array_equal(sdp, dp[ri, ci]) and np.array_equal(sal, al[ri, ci])
    q = sal > 0
    if skip is not None:
        q &= skip[site] == 0
    assert np.all(np.isnan(p[~q]))
    ref = betabinom.sf(sal[q] - 0.001, sdp[q], a2, b2)
    assert np.all(np.abs(p[q] - ref) <= 1e-9 * ref + 1e-12)
    assert np.array_equal(np.round(p[q], 4), np.round(ref, 4))
    return int(q.sum())


def test_sparse_genotyping_equals_dense(engine):
    import oracle
    cfg = synth.config("C5", scale=0.02)
    cfg["n_cells"] = 2000
    d = synth.generate(**cfg)
    sc = engine.pileup_count(d.batch, _windows(d), CountParams(min_mq=60))
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(sc.n_sites, size=min(5000, sc.n_sites), replace=False))
    alt = rng.integers(0, 6, size=idx.shape[0]).astype(np.uint8)
    a2, b2 = 0.2474528917555431, 162.03696139428595
    odp, oal = oracle.genotype_count(d.batch, sc.tid[idx], sc.pos[idx], alt, 2000, min_bq=30, min_mq=60)
    assert _check_sparse_against_dense(engine, sc.tid[idx], sc.pos[idx], alt, 2000, odp, oal, a2, b2) > 100
    skip = (np.arange(idx.shape[0]) % 3 == 0).astype(np.uint8)  # every third site takes the chrM shortcut: no tail
    _check_sparse_against_dense(engine, sc.tid[idx], sc.pos[idx], alt, 2000, odp, oal, a2, b2, skip=skip)
    # --alt_flag Alt
    odp, oal = oracle.genotype_count(d.batch, sc.tid[idx], sc.pos[idx], alt, 2000, min_bq=30, min_mq=60, alt_only=True)
    site, cell, sdp, sal, p = engine.genotype_sparse(sc.tid[idx], sc.pos[idx], alt, 2000, a2, b2, min_bq=30, min_mq=60,
                                                     alt_only=True)
    ri, ci = np.nonzero(odp > 0)
    assert np.array_equal(site, ri) and np.array_equal(cell, ci) and np.array_equal(sdp, odp[ri, ci])
    assert np.array_equal(sal, oal[ri, ci])


def test_c5_top_of_grid_sparse(engine):
    """BASELINE.json configs[4] at the top of its grid: 200 000 candidate sites x 20 000 cells (dense tensors would be
    32 GB).  The sparse tuples are compared with the oracle's dense tensors site chunk by site chunk."""
    import oracle
    d = synth.generate(**synth.config("C5", scale=1.0))
    sc = engine.pileup_count(d.batch, _windows(d), CountParams(min_mq=60))
    n_cells = 20000
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(sc.n_sites, size=200000, replace=False))
    alt = rng.integers(0, 4, size=idx.shape[0]).astype(np.uint8)
    a2, b2 = 0.2474528917555431, 162.03696139428595
    site, cell, sdp, sal, p = engine.genotype_sparse(sc.tid[idx], sc.pos[idx], alt, n_cells, a2, b2, min_bq=30, min_mq=60)
    assert engine.last_stats["ms_total"] < 10000.0
    step = 10000
    for c0 in range(0, idx.shape[0], step * 4):  # every fourth chunk: 5 x (10 000 x 20 000) dense oracle tensors
        j = idx[c0:c0 + step]
        odp, oal = oracle.genotype_count(d.batch, sc.tid[j], sc.pos[j], alt[c0:c0 + step], n_cells, min_bq=30, min_mq=60)
        ri, ci = np.nonzero(odp > 0)
        lo, hi = np.searchsorted(site, c0), np.searchsorted(site, c0 + len(j))
        assert np.array_equal(site[lo:hi] - c0, ri) and np.array_equal(cell[lo:hi], ci)
        assert np.array_equal(sdp[lo:hi], odp[ri, ci]) and np.array_equal(sal[lo:hi], oal[ri, ci])
    q = sal > 0
    assert q.sum() > 0 and np.all(np.isfinite(p[q])) and np.all((p[q] >= 0) & (p[q] <= 1)) and np.all(np.isnan(p[~q]))


def test_linearity_over_disjoint_cell_sets(engine):
    """Size-independent property on a larger batch (250k reads): every output word is additive over
    batches with disjoint cells -- counts(all) == counts(even cells) + counts(odd cells)."""
    d = synth.generate(**synth.config("C2", scale=0.05))
    w = _windows(d)
    p = CountParams(min_bq=20, min_mq=60, min_dp=1, min_cc=1)
    b = d.batch

    def sub(mask):
        cell = np.where(mask, b.cell, -1).astype(np.int32)  # other reads lose their CB: seen but never counted
        return ReadBatch(b.tid, b.pos, b.flag, b.mapq, cell, b.cigar_off, b.cigar, b.base_off, b.l_qseq, b.seq4, b.qual)
    full = engine.pileup_count(b, w, p)
    even = engine.pileup_count(sub((b.cell >= 0) & (b.cell % 2 == 0)), w, p)
    odd = engine.pileup_count(sub((b.cell >= 0) & (b.cell % 2 == 1)), w, p)
    key = lambda s: (s.tid.astype(np.int64) << 32) | s.pos.astype(np.int64)
    acc = np.zeros_like(full.counts, dtype=np.int64)
    kf = key(full)
    for part in (even, odd):
        j = np.searchsorted(kf, key(part))
        assert np.array_equal(kf[j], key(part))
        acc[j] += part.counts
    assert np.array_equal(acc, full.counts.astype(np.int64))
    # and the whole table survives a (tile, cell) re-numbering: permuting cell ids changes nothing
    perm = np.random.default_rng(0).permutation(int(b.cell.max()) + 1).astype(np.int32)
    pb = ReadBatch(b.tid, b.pos, b.flag, b.mapq, np.where(b.cell >= 0, perm[np.maximum(b.cell, 0)], -1).astype(np.int32),
                   b.cigar_off, b.cigar, b.base_off, b.l_qseq, b.seq4, b.qual)
    again = engine.pileup_count(pb, w, p)
    assert np.array_equal(again.counts, full.counts)
