"""CPU suite: pins the oracle and the host logic to the REFERENCE's own outputs.

tests/golden/<case>/*.gz were produced by the unmodified reference scripts (oracle/make_golden.py,
run in the build container over oracle/shims).  Here, without a GPU:
  * the C oracle (oracle/pileup_oracle.c) must reproduce the reference's BaseCellCounter tables
    byte for byte  -> the oracle is pinned;
  * the host side of every drop-in CLI (parsing, label cascades, row order, formatting) must
    reproduce the reference's merged / step1 / step2 / genotype files byte for byte when the GPU
    calls are answered by the oracle (C pileup, scipy beta-binomial, numpy set membership).
The GPU suite (test_cli_gpu.py) then runs the same CLIs with the real CUDA engine."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = [c for c in ("g1", "g2", "g3") if os.path.isdir(os.path.join(GOLD, c))]


def gold_lines(case, name):
    if not os.path.exists(os.path.join(GOLD, case, name + ".gz")):
        return HashedGold(os.path.join(GOLD, case, name + ".sha256.json"))
    with gzip.open(os.path.join(GOLD, case, name + ".gz"), "rt") as f:
        return [l for l in f if not l.startswith("##fileDate")]


def file_lines(path):
    with open(path) as f:
        return [l for l in f if not l.startswith("##fileDate")]


def has_gold(case, name):
    return os.path.exists(os.path.join(GOLD, case, name + ".gz")) or os.path.exists(os.path.join(GOLD, case, name + ".sha256.json"))


class HashedGold:
    """A golden too large to commit (g3's 176 MB per-cell genotype table): line count + SHA-256 of its text."""

    def __init__(self, path):
        import json
        meta = json.load(open(path))
        self.sha256, self.lines = meta["sha256"], meta["lines"]


def assert_same(got, want, what):
    if isinstance(want, HashedGold):
        import hashlib
        assert len(got) == want.lines, "%s: %d lines vs %d" % (what, len(got), want.lines)
        assert hashlib.sha256("".join(got).encode()).hexdigest() == want.sha256, what + ": SHA-256 differs"
        return
    assert len(got) == len(want), "%s: %d lines vs %d" % (what, len(got), len(want))
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, "%s differs at line %d:\n got: %s\nwant: %s" % (what, i, a[:400], b[:400])


@pytest.fixture(scope="module", params=CASES)
def work(request, tmp_path_factory, built):
    import pipeline_inputs as pi
    case = request.param
    d = tmp_path_factory.mktemp("in_" + case)
    paths, data = pi.write_inputs(case, str(d))
    # golden intermediate files, so that each stage is tested in isolation
    for name in ("counts.Cancer.tsv", "counts.Non-Cancer.tsv", "merged.tsv", "step1.tsv", "candidates.tsv"):
        with gzip.open(os.path.join(GOLD, case, name + ".gz"), "rb") as f, open(os.path.join(str(d), name), "wb") as o:
            o.write(f.read())
    return case, str(d), paths, data


class OracleEngine:
    """Answers the Engine calls of the CLIs with the CPU oracle (test infrastructure)."""

    def __init__(self, *a, **k):
        self.batch = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass

    last_stats = {}

    def upload(self, batch, windows=None):
        self.batch = batch

    def pileup_count(self, batch, windows, params):
        import oracle
        return oracle.pileup_count(batch, windows, params, threads=4)[0]

    def genotype_count(self, st, sp, ac, n_cells, **kw):
        import oracle
        return oracle.genotype_count(self.batch, st, sp, ac, n_cells, **kw)

    def betabinom_sf(self, k, n, a, b):
        from scipy.stats import betabinom
        k, n = np.asarray(k), np.asarray(n)
        return betabinom.sf(k - 0.1, n, a, b) if len(k) else np.zeros(0)

    def site_mask(self, keys, query):
        return np.isin(query, keys).astype(np.uint8)


def _count_with_oracle(bam, ref, chrom, bin_size, prm, ID, out, bed="", bed_out=""):
    from longsom_b200 import bamio
    from longsom_b200.batch import Windows
    from longsom_b200.pipeline import load_bam_for_counting, prune_and_sort_windows, read_ends, write_counter_tsv
    from longsom_b200.windows import make_windows
    import oracle
    fa = bamio.Fasta(ref)
    named = make_windows(fa.references, fa.lengths, chrom, bin_size, bed=bed, bed_out=bed_out)
    bd, batch, _ = load_bam_for_counting(bam)
    iv = prune_and_sort_windows(named, bd.contig_names, batch, read_ends(batch))
    win = Windows.from_intervals(iv, {t: fa.contig(bd.contig_names[t]) for t in {w[0] for w in iv}})
    sc, _ = oracle.pileup_count(batch, win, prm, threads=4)
    write_counter_tsv(out, ID, sc, bd.contig_names)


def test_oracle_pileup_matches_reference_tables(work):
    from longsom_b200.engine import CountParams
    case, d, p, data = work
    for name, bam in (("Cancer", p["cancer"]), ("Non-Cancer", p["normal"])):
        out = os.path.join(d, "o_%s.tsv" % name)
        _count_with_oracle(bam, p["ref"], "all", 50000, CountParams(min_bq=20, min_mq=60), "s." + name, out)
        assert_same(file_lines(out), gold_lines(case, "counts.%s.tsv" % name), "BaseCellCounter " + name)
    if not has_gold(case, "counts.full_ac.tsv"):
        return  # core-stages-only case (g3)
    out = os.path.join(d, "o_ac.tsv")
    _count_with_oracle(p["full"], p["ref"], data.contig_names[0], 30000,
                       CountParams(min_bq=30, min_mq=0, min_ac=2, min_dp=3, min_cc=2), "full.ac", out)
    assert_same(file_lines(out), gold_lines(case, "counts.full_ac.tsv"), "BaseCellCounter --min_ac 2")


class _NumpySlot:
    """Stand-in for pipeline.PinnedSlot on a box without a GPU (pinned memory needs a CUDA context)."""

    def __call__(self, n, nc, nb):
        return dict(tid=np.zeros(n, np.int32), pos=np.zeros(n, np.int32), flag=np.zeros(n, np.uint16),
                    mapq=np.zeros(n, np.uint8), cell=np.zeros(n, np.int32), l_qseq=np.zeros(n, np.int32),
                    cigar_off=np.zeros(n + 1, np.uint32), base_off=np.zeros(n + 1, np.uint64), cigar=np.zeros(nc, np.uint32),
                    seq4=np.zeros(nb // 2 + 1, np.uint8), qual=np.zeros(nb, np.uint8))

    def free(self):
        pass


@pytest.mark.parametrize("chunk_kb", [64, 700])
def test_streaming_counter_host_logic_matches_reference(work, monkeypatch, chunk_kb):
    """The streaming BaseCellCounter (chunked BAM decode, reads carried across chunks, windows completed on the fly,
    per-contig row files merged in name order) with the device calls answered by the oracle: byte-identical tables,
    with chunks far smaller than a 50 kb window's reads (64 KB) and with a few chunks per contig (700 KB)."""
    from longsom_b200 import bamio, pipeline
    from longsom_b200.engine import CountParams
    from longsom_b200.windows import make_windows
    case, d, p, data = work
    monkeypatch.setattr(pipeline, "PinnedSlot", _NumpySlot)
    monkeypatch.setattr(pipeline, "take_engine", lambda dev: OracleEngine())
    fa = bamio.Fasta(p["ref"])
    named = make_windows(fa.references, fa.lengths, "all", 50000)
    for name, bam in (("Cancer", p["cancer"]), ("Non-Cancer", p["normal"])):
        out = os.path.join(d, "stream_%s_%d.tsv" % (name, chunk_kb))
        n = pipeline.stream_count(bam, named, fa, CountParams(min_bq=20, min_mq=60), out, "s." + name, [0],
                                  chunk_bytes=chunk_kb * 1024)
        assert n > 0
        assert_same(file_lines(out), gold_lines(case, "counts.%s.tsv" % name), "streaming BaseCellCounter " + name)
        assert not [f for f in os.listdir(d) if ".part." in f]  # the per-contig row files are removed


def test_bed_and_bed_out_windows_match_reference(work):
    """MakeWindows with --bed / --bed_out (BaseCellCounter.py:87-110): the host interval arithmetic of
    longsom_b200/windows.py against tables the reference wrote over the pybedtools stand-in (golden g1)."""
    import pipeline_inputs as pi
    from longsom_b200.engine import CountParams
    case, d, p, data = work
    if not has_gold(case, "counts.full_bed.tsv"):
        return  # the --bed goldens were generated for g1
    bed, bed_out = pi.write_beds(case, d, data)
    prm = CountParams(min_bq=20, min_mq=60, min_dp=3, min_cc=2)
    for name, kw, chrom, bin_size in (("bed", dict(bed=bed), "all", 50000),
                                      ("bed_out", dict(bed=bed, bed_out=bed_out), data.contig_names[0], 50000),
                                      ("bedout_only", dict(bed_out=bed_out), "all", 20000)):
        out = os.path.join(d, "o_%s.tsv" % name)
        _count_with_oracle(p["full"], p["ref"], chrom, bin_size, prm, "full." + name, out, **kw)
        assert_same(file_lines(out), gold_lines(case, "counts.full_%s.tsv" % name), "BaseCellCounter --" + name)


def test_windows_with_1e5_intervals_are_fast():
    """bedtools does these as sorted sweeps; the stand-alone arithmetic must not be quadratic."""
    import time
    from longsom_b200 import windows as W
    rng = np.random.default_rng(3)
    starts = np.sort(rng.integers(0, 50_000_000, size=100_000))
    a = [("chr1", int(s), int(s) + 120) for s in starts]
    b = [("chr1", int(s) + 60, int(s) + 90) for s in starts[::2]] + [("chr2", 5, 10)]
    t0 = time.time()
    m = W._merge(a, 1)
    i = W._intersect(m, [("chr1", 1, 40_000_000), ("chr2", 1, 1000)])
    r = W._subtract(i, b)
    assert time.time() - t0 < 20.0
    assert len(r) >= len(i) and all(e > s for _, s, e in r)


def test_merge_matches_reference(work):
    from longsom_b200.cli.merge import merge_cell_types_files
    case, d, p, data = work
    want = gold_lines(case, "merged.tsv")
    order = want[7].rstrip("\n").split("\t")[5:]  # the reference's glob order decides the column order (Q10)
    folder = os.path.join(d, "mc")
    os.makedirs(folder, exist_ok=True)
    files = []
    for t in order:
        fp = os.path.join(folder, "s.%s.tsv" % t)
        with gzip.open(os.path.join(GOLD, case, "counts.%s.tsv.gz" % t), "rb") as f, open(fp, "wb") as o:
            o.write(f.read())
        files.append(fp)
    out = os.path.join(d, "merged_mine.tsv")
    merge_cell_types_files(files, out)
    assert_same(file_lines(out), want, "MergeBaseCellCounts")


def test_merge_cursor_quirks(tmp_path):
    """The lock-step cursors of MergeBaseCellCounts.py:8-23,116-204 on crafted tables: a repeated position and a
    position that goes backwards are skipped, a blank line ends its table, chromosomes come in lexicographic order
    (chr10 before chr2), absent cell types read 'NA'.  The expected rows are the reference's own output on these
    inputs (generated once with the reference script in the build container)."""
    from longsom_b200.cli.merge import merge_cell_types_files
    fmt = "DP|NC|CC|BC|BQ|BCf|BCr"
    hdr = "##fileDate=x\n" + "\n".join("##INFO=%d" % i for i in range(7)) + "\n#CHROM\tStart\tREF\tINFO\tX\n"
    a = [("chr1", 5, "A", "a5"), ("chr1", 9, "C", "a9"), ("chr1", 9, "C", "a9dup"), ("chr1", 7, "G", "a7back"),
         ("chr10", 3, "T", "a10_3"), ("chr2", 1, "G", "a2_1")]
    b = [("chr1", 9, "c", "b9"), ("chr2", 1, "G", "b2_1"), ("chr2", 8, ".", "b2_8")]
    c = [("chr10", 3, "T", "c10_3")]
    files = []
    for name, rows, tail in (("s.A.tsv", a, ""), ("s.B.tsv", b, "\nchr3\t1\tA\t%s\tafter_blank\n" % fmt), ("s.C.tsv", c, "")):
        fp = str(tmp_path / name)
        with open(fp, "w") as f:
            f.write(hdr)
            for ch, pos, ref, bc in rows:
                f.write("%s\t%d\t%s\t%s\t%s\n" % (ch, pos, ref, fmt, bc))
            f.write(tail)
        files.append(fp)
    out = str(tmp_path / "merged.tsv")
    merge_cell_types_files(files, out)
    got = open(out).read().splitlines()
    assert got[8] == "#CHROM\tStart\tEnd\tREF\tINFO\tA\tB\tC"
    assert got[9:] == ["chr1\t5\t5\tA\t%s\ta5\tNA\tNA" % fmt,
                       "chr1\t9\t9\tC|c\t%s\ta9\tb9\tNA" % fmt,
                       "chr10\t3\t3\tT\t%s\ta10_3\tNA\tc10_3" % fmt,
                       "chr2\t1\t1\tG\t%s\ta2_1\tb2_1\tNA" % fmt,
                       "chr2\t8\t8\t.\t%s\tNA\tb2_8\tNA" % fmt]


def test_step1_host_logic_matches_reference(work):
    import pipeline_inputs as pi
    from longsom_b200.cli.step1 import variant_calling_step1
    case, d, p, data = work
    out = os.path.join(d, "step1_mine.tsv")
    variant_calling_step1(os.path.join(d, "merged.tsv"), out, p["ref"], pi.ALPHA1, pi.BETA1, pi.ALPHA2, pi.BETA2, 2, 3, 5, 5,
                          2, 1, 1, OracleEngine())
    assert_same(file_lines(out), gold_lines(case, "step1.tsv"), "BaseCellCalling.step1")


def test_step1_line_handling_follows_text_mode(work):
    """The reference reads its table in text mode (BaseCellCalling.step1.py:32-35): '\\r\\n' ends a line like '\\n', and
    a '##' line in the BODY is copied through where it stands.  Same table with CRLF line ends and a comment in the
    middle -> the golden output plus that comment, in the single-process and in the forked-workers path."""
    import pipeline_inputs as pi
    from longsom_b200.cli.step1 import variant_calling_step1
    case, d, p, data = work
    src = open(os.path.join(d, "merged.tsv")).read().split("\n")
    n_head = next(i for i, x in enumerate(src) if x.startswith("#CHROM")) + 1
    body = [x for x in src[n_head:] if x]
    k = len(body) // 2
    crlf = os.path.join(d, "merged_crlf.tsv")
    with open(crlf, "w", newline="") as f:
        f.write("\r\n".join(src[:n_head] + body[:k] + ["##a comment in the body"] + body[k:]) + "\r\n")
    want = gold_lines(case, "step1.tsv")
    w_head = next(i for i, x in enumerate(want) if x.startswith("#CHROM")) + 1
    want = want[:w_head + k] + ["##a comment in the body\n"] + want[w_head + k:]
    for procs in (1, 3):
        out = os.path.join(d, "step1_crlf_%d.tsv" % procs)
        variant_calling_step1(crlf, out, p["ref"], pi.ALPHA1, pi.BETA1, pi.ALPHA2, pi.BETA2, 2, 3, 5, 5, 2, 1, 1,
                              OracleEngine(), procs=procs)
        assert_same(file_lines(out), want, "BaseCellCalling.step1 (CRLF, procs=%d)" % procs)


def test_step2_host_logic_matches_reference(work):
    from longsom_b200.cli.step2 import variant_calling_step2
    case, d, p, data = work
    out = os.path.join(d, "step2_mine.tsv")
    variant_calling_step2(os.path.join(d, "step1.tsv"), 0, p["editing"], p["pon_sr"], p["pon_lr"], p["gnomad"], 0.01, out,
                          OracleEngine())
    assert_same(file_lines(out), gold_lines(case, "step2.tsv"), "BaseCellCalling.step2")
    if not has_gold(case, "step2_gz.tsv"):
        return
    out = os.path.join(d, "step2gz_mine.tsv")
    variant_calling_step2(os.path.join(d, "step1.tsv"), 5, p["editing_gz"], p["pon_sr"], "", p["gnomad"], 0.01, out,
                          OracleEngine())
    assert_same(file_lines(out), gold_lines(case, "step2_gz.tsv"), "BaseCellCalling.step2 (gz editing list)")


class _GenoOracleEngine(OracleEngine):
    def betabinom_sf(self, k, n, a, b):
        from scipy.stats import betabinom
        k, n = np.asarray(k), np.asarray(n)
        return betabinom.sf(k - 0.001, n, a, b) if len(k) else np.zeros(0)

    def genotype_sparse(self, st, sp, ac, n_cells, alpha, beta, skip_p=None, **kw):
        """The sparse tuples of Engine.genotype_sparse, from the oracle's dense tensors and scipy."""
        dp, alt = self.genotype_count(st, sp, ac, n_cells, **kw)
        ri, ci = np.nonzero(dp > 0)
        d, a = dp[ri, ci].astype(np.int32), alt[ri, ci].astype(np.int32)
        q = a > 0
        if skip_p is not None:
            q &= np.asarray(skip_p)[ri] == 0
        p = np.full(len(ri), np.nan)
        p[q] = self.betabinom_sf(a[q], d[q], alpha, beta)
        return ri.astype(np.int32), ci.astype(np.int32), d, a, p


def test_genotype_host_logic_matches_reference(work, monkeypatch):
    import pipeline_inputs as pi
    import longsom_b200.cli.genotype as G
    case, d, p, data = work
    monkeypatch.setattr(G, "Engine", _GenoOracleEngine)
    for flag in ("All", "Alt"):
        if not has_gold(case, "geno_%s.DpMatrix.tsv" % flag):
            continue
        pre = os.path.join(d, "geno_" + flag)
        G.main(["--bam", p["full"], "--infile", os.path.join(d, "candidates.tsv"), "--ref", p["ref"], "--meta", p["meta"],
                "--fusions", "--outfile", pre, "--alt_flag", flag, "--min_mq", "60", "--alpha2", str(pi.ALPHA2), "--beta2",
                str(pi.BETA2), "--tmp_dir", os.path.join(d, "tmpg")])
        for suf in ("SingleCellGenotype", "DpMatrix", "AltMatrix", "VAFMatrix", "BinaryMatrix"):
            assert_same(file_lines("%s.%s.tsv" % (pre, suf)), gold_lines(case, "geno_%s.%s.tsv" % (flag, suf)),
                        "SingleCellGenotype %s %s" % (flag, suf))
    if not has_gold(case, "hccv.tsv"):
        return
    out = os.path.join(d, "hccv_mine.tsv")
    G.main(["--bam", p["full"], "--infile", os.path.join(d, "candidates.tsv"), "--ref", p["ref"], "--meta", p["meta"],
            "--outfile", out, "--alt_flag", "All", "--min_mq", "60", "--tmp_dir", os.path.join(d, "tmph")], hccv=True)
    assert_same(file_lines(out), gold_lines(case, "hccv.tsv"), "HCCVSingleCellGenotype")


def test_shim_engine_matches_c_oracle_on_corner_cases(built):
    """The literal htslib-engine restatement (oracle/shims/pysam) and the C oracle agree on the
    hand-built CIGAR corner cases, independently of the reference scripts."""
    import pipeline_inputs as pi
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    import importlib
    pysam = importlib.import_module("pysam")
    from longsom_b200 import bamio
    from longsom_b200.batch import Windows
    from longsom_b200.engine import CountParams
    import oracle
    import tempfile
    d = pi.build_batch("g2")
    tmp = tempfile.mkdtemp()
    bam = os.path.join(tmp, "x.bam")
    bamio.write_bam(bam, d.contig_names, d.contig_lens, d.batch, pi.cb_text(d, ""))
    lo, hi = 250, 520
    win = Windows.from_intervals([(0, lo, hi)], d.contig_seqs())
    prm = CountParams(min_bq=20, min_mq=60, min_dp=1, min_cc=1)
    sc, _ = oracle.pileup_count(d.batch, win, prm)
    dp_oracle = dict(zip(sc.pos.tolist(), sc.counts[:, 0].tolist()))
    a = pysam.AlignmentFile(bam)
    dp_shim = {}
    for col in a.pileup(d.contig_names[0], lo, hi, min_base_quality=20, min_mapping_quality=60, ignore_overlaps=False,
                        max_depth=200000):
        if not (lo <= col.pos < hi):
            continue
        n = 0
        for s, pr in zip(col.get_query_sequences(mark_matches=True, add_indels=True), col.pileups):
            al = pr.alignment
            if "CB" not in al.tags or al.is_supplementary:
                continue
            u = s.upper()
            if u in ("A", "C", "G", "T", "N") or (len(s) > 1 and s[1] in "+-") or s == "*":
                n += 1
        if n and chr(d.contig_seq(0)[col.pos]).upper() != "N":
            dp_shim[col.pos] = n
    assert dp_shim == dp_oracle
    assert len(dp_shim) > 50


STEP3_SETS = {"g1": (("step2.tsv", "step3", 0.05, 0.3, 3, 2, 10000), ("step2.tsv", "step3_loose", 0.05, 0.05, 1, 1, 50)),
              "g2": (("step2.tsv", "step3", 0.05, 0.3, 3, 2, 10000), ("step2.tsv", "step3_loose", 0.05, 0.05, 1, 1, 50)),
              "s3": (("step2_fabricated.tsv", "step3", 0.2, 0.25, 3, 2, 1000),)}


@pytest.mark.parametrize("case", [c for c in STEP3_SETS if os.path.isdir(os.path.join(GOLD, c))])
def test_step3_matches_reference(case, tmp_path):
    """the step3 drop-in (host only) against the reference script's two output files; case s3
    is the fabricated table that reaches the chrM / multi-allelic / cluster branches."""
    from longsom_b200.cli.step3 import main
    for table, name, dvaf, dmcf, mr, mc, cd in STEP3_SETS[case]:
        src = os.path.join(str(tmp_path), table)
        with gzip.open(os.path.join(GOLD, case, table + ".gz"), "rb") as f, open(src, "wb") as o:
            o.write(f.read())
        pre = os.path.join(str(tmp_path), name)
        main(["--infile", src, "--outfile", pre, "--deltaVAF", str(dvaf), "--deltaMCF", str(dmcf), "--min_ac_reads", str(mr),
              "--min_ac_cells", str(mc), "--clust_dist", str(cd)])
        assert_same(file_lines(pre + ".calling.step3.tsv"), gold_lines(case, name + ".tsv"), "step3 " + name)
        assert_same(file_lines(pre + ".calling.step3.unfiltered.tsv"), gold_lines(case, name + ".unfiltered.tsv"),
                    "step3 unfiltered " + name)


def test_step3_wrapper_script_runs(tmp_path):
    import subprocess
    src = os.path.join(str(tmp_path), "in.tsv")
    with gzip.open(os.path.join(GOLD, "s3", "step2_fabricated.tsv.gz"), "rb") as f, open(src, "wb") as o:
        o.write(f.read())
    pre = os.path.join(str(tmp_path), "w")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "workflow", "scripts", "SNVCalling", "BaseCellCalling.step3.py"),
                        "--infile", src, "--outfile", pre, "--deltaVAF", "0.2", "--deltaMCF", "0.25", "--min_ac_reads", "3",
                        "--min_ac_cells", "2", "--clust_dist", "1000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert_same(file_lines(pre + ".calling.step3.tsv"), gold_lines("s3", "step3.tsv"), "step3 wrapper")


HCCV_SETS = (("hccv_variants", 20, 0.1, 0.4, 10000), ("hccv_variants_loose", 5, 0.05, 0.05, 30))


@pytest.mark.parametrize("case", [c for c in STEP3_SETS if os.path.isdir(os.path.join(GOLD, c))])
def test_hccv_variants_match_reference(case, tmp_path):
    """HighConfidenceCancerVariants drop-in (host only): the final table and the two intermediate
    dumps (.tsv2 after the depth filter, .tsv3 with the verdict column) against the reference's."""
    from longsom_b200.cli.hccv_variants import main
    table = STEP3_SETS[case][0][0]
    src = os.path.join(str(tmp_path), table)
    with gzip.open(os.path.join(GOLD, case, table + ".gz"), "rb") as f, open(src, "wb") as o:
        o.write(f.read())
    for name, min_dp, dvaf, dmcf, cd in HCCV_SETS:
        pre = os.path.join(str(tmp_path), name)
        argv = ["--SNVs", src, "--outfile", pre, "--min_dp", str(min_dp), "--deltaVAF", str(dvaf), "--deltaMCF", str(dmcf),
                "--clust_dist", str(cd)]
        if not os.path.exists(os.path.join(GOLD, case, name + ".tsv.gz")):
            # the reference dies inside pandas when no variant survives (g2 with the default thresholds);
            # the drop-in must not turn that into a silent empty success
            with pytest.raises((AttributeError, ValueError, KeyError)):
                main(argv)
            continue
        main(argv)
        for suf in ("", "2", "3"):
            assert_same(file_lines(pre + ".HCCV.tsv" + suf), gold_lines(case, name + ".tsv" + suf), "HCCV %s%s" % (name, suf))


def test_reannotation_matches_reference(work):
    """CellTypeReannotation drop-in on the golden HCCV genotype table (K1' output), the case's barcode
    table and the committed fabricated fusion table; three (min_variants, min_frac) settings."""
    from longsom_b200.cli.reannotate import main
    case, d, p, data = work
    if not has_gold(case, "hccv.tsv"):
        return  # core-stages-only case (g3)
    for name in ("hccv.tsv", "reannot_fusions.tsv"):
        with gzip.open(os.path.join(GOLD, case, name + ".gz"), "rb") as f, open(os.path.join(d, name), "wb") as o:
            o.write(f.read())
    for name, mv, mf in (("reannot", 3, 0.2), ("reannot_loose", 1, 0.02), ("reannot_mid", 5, 0.03)):
        out = os.path.join(d, name + "_mine.tsv")
        main(["--SNVs", os.path.join(d, "hccv.tsv"), "--fusions", os.path.join(d, "reannot_fusions.tsv"), "--outfile", out,
              "--meta", p["meta"], "--min_variants", str(mv), "--min_frac", str(mf)])
        assert_same(file_lines(out), gold_lines(case, name + ".tsv"), "CellTypeReannotation " + name)


def test_filters_fail_like_reference_on_empty_selection(tmp_path):
    """Only 'Non-Cancer' rows: the reference's pandas calls raise ValueError after the header has been written."""
    from longsom_b200.cli import hccv_variants, step3
    src = os.path.join(str(tmp_path), "in.tsv")
    with gzip.open(os.path.join(GOLD, "s3", "step2_fabricated.tsv.gz"), "rt") as f, open(src, "w") as o:
        for line in f:
            if line.startswith("#") or line.split("\t")[6] == "Non-Cancer":
                o.write(line)
    with pytest.raises(ValueError):
        step3.main(["--infile", src, "--outfile", os.path.join(str(tmp_path), "e"), "--deltaVAF", "0.1", "--deltaMCF", "0.3"])
    with pytest.raises(ValueError):
        hccv_variants.main(["--SNVs", src, "--outfile", os.path.join(str(tmp_path), "e"), "--min_dp", "20", "--deltaVAF", "0.1",
                            "--deltaMCF", "0.3"])
    assert os.path.exists(os.path.join(str(tmp_path), "e.calling.step3.tsv"))


def test_step1_worker_processes_do_not_change_the_output(work):
    """step1 cuts large tables into byte ranges handled by forked workers (queries gathered in the parent, tails
    sent back, lines concatenated in range order): forced here on the golden table, 1..5 workers."""
    import pipeline_inputs as pi
    from longsom_b200.cli.step1 import variant_calling_step1
    case, d, p, data = work
    for procs in ((3,) if case == "g3" else (2, 3, 5)):
        out = os.path.join(d, "step1_mp%d.tsv" % procs)
        n_rows, n_q = variant_calling_step1(os.path.join(d, "merged.tsv"), out, p["ref"], pi.ALPHA1, pi.BETA1, pi.ALPHA2, pi.BETA2,
                                            2, 3, 5, 5, 2, 1, 1, OracleEngine(), procs=procs)
        assert_same(file_lines(out), gold_lines(case, "step1.tsv"), "BaseCellCalling.step1 with %d workers" % procs)
        assert n_rows == len([l for l in gold_lines(case, "step1.tsv") if not l.startswith("#")])
