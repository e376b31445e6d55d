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


@pytest.mark.parametrize("n_sites,n_cells", [(1000, 1000), (10000, 1000), (3000, 5000)])
def test_c5_genotyping_sweep(engine, n_sites, n_cells):
    import oracle
    from scipy.stats import betabinom
    cfg = synth.config("C5", scale=0.02)
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
