"""the SplitBamCellTypes drop-in (native record router + BGZF writer + .bai) against goldens made by
the unmodified reference script over the pysam shim.  Compared per output BAM: header bytes, every record byte for
byte (md5 of the raw record, qualities shown in clear), record order; the report without its run-time column.
The index is checked functionally: samtools-style region queries through the .bai equal a brute-force scan."""
import gzip
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
GOLD = os.path.join(ROOT, "tests", "golden")
SETS = (("split", ["--min_MQ", "60", "--n_trim", "5"]),
        ("split_all", ["--min_MQ", "30", "--n_trim", "0", "--max_nM", "5", "--max_NH", "1"]))


def gold(case, name):
    with gzip.open(os.path.join(GOLD, case, name + ".gz"), "rt") as f:
        return f.read().splitlines()


@pytest.fixture(scope="module", params=[c for c in ("g1", "g2") if os.path.exists(os.path.join(GOLD, c, "split.report.txt.gz"))])
def split_runs(request, tmp_path_factory, built):
    import pipeline_inputs as pi
    from longsom_b200.cli import splitbam
    case = request.param
    d = str(tmp_path_factory.mktemp("split_" + case))
    bam, meta = pi.write_split_input(case, d)
    outs = {}
    for name, extra in SETS:
        out = os.path.join(d, name)
        os.makedirs(out)
        splitbam.main(["--bam", bam, "--meta", meta, "--id", "s", "--outdir", out] + extra)
        outs[name] = out
    return case, d, bam, meta, outs


def test_split_records_and_report_match_reference(split_runs):
    import pipeline_inputs as pi
    case, d, bam, meta, outs = split_runs
    for name, out in outs.items():
        bams = sorted(f for f in os.listdir(out) if f.endswith(".bam"))
        want = sorted(f[len(name) + 1:-len(".records.txt.gz")] + ".bam" for f in os.listdir(os.path.join(GOLD, case))
                      if f.startswith(name + ".s.") and f.endswith(".records.txt.gz"))
        assert bams == want
        for fn in bams:
            got = pi.dump_bam_records(os.path.join(out, fn))
            ref = gold(case, "%s.%s.records.txt" % (name, fn[:-4]))
            assert len(got) == len(ref), (name, fn, len(got), len(ref))
            for i, (a, b) in enumerate(zip(got, ref)):
                assert a == b, "%s %s record %d:\n got %s\nwant %s" % (name, fn, i, a[:300], b[:300])
            assert os.path.getsize(os.path.join(out, fn + ".bai")) > 8
            assert open(os.path.join(out, fn), "rb").read()[-28:] == bytes.fromhex(
                "1f8b08040000000000ff0600424302001b0003000000000000000000")   # BGZF EOF marker
        rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(out, "s.report.txt"))]
        assert rows[0][-1] == "Total_time"
        assert ["\t".join(r[:-1]) for r in rows] == gold(case, name + ".report.txt")


def test_split_outputs_decode_with_the_product_reader(split_runs):
    """The outputs are what BaseCellCounter reads next: the native decoder must accept them."""
    from longsom_b200 import bamio
    case, d, bam, meta, outs = split_runs
    total = 0
    for fn in os.listdir(outs["split"]):
        if fn.endswith(".bam"):
            total += bamio.read_bam(os.path.join(outs["split"], fn)).batch.n_reads
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(outs["split"], "s.report.txt"))]
    assert total == int(rows[1][rows[0].index("Pass_reads")])


def test_bai_region_queries_equal_brute_force(split_runs):
    import bai_query
    import pipeline_inputs as pi
    case, d, bam, meta, outs = split_runs
    rng = random.Random(5)
    checked = 0
    for fn in sorted(os.listdir(outs["split"])):
        if not fn.endswith(".bam"):
            continue
        path = os.path.join(outs["split"], fn)
        recs = [l.split("\t") for l in pi.dump_bam_records(path)[1:]]
        refs, n_no_coor = bai_query.read_bai(path + ".bai")
        assert n_no_coor == 0
        # brute force needs the reference span: take it from the product decoder
        from longsom_b200 import bamio
        from longsom_b200.pipeline import read_ends
        bd = bamio.read_bam(path)
        ends = read_ends(bd.batch)
        tids, poss = bd.batch.tid, bd.batch.pos
        names = [r[0] for r in recs]
        assert len(names) == bd.batch.n_reads
        n_tid = int(tids.max()) + 1 if len(tids) else 0
        for _ in range(25):
            t = rng.randrange(max(n_tid, 1))
            sel = [i for i in range(len(names)) if tids[i] == t]
            if not sel:
                continue
            lo = rng.choice(sel)
            beg = max(0, int(poss[lo]) + rng.randrange(-50, 50))
            end = beg + rng.choice([1, 10, 300, 20000, 100000])
            want = sorted(set((names[i], int(poss[i])) for i in sel
                              if poss[i] < end and max(int(ends[i]), int(poss[i]) + 1) > beg))
            assert bai_query.query(path, t, beg, end) == want, (fn, t, beg, end)
            checked += 1
    assert checked > 20


def test_split_fails_loudly_like_the_reference(tmp_path, built):
    """A trim longer than a read is an IndexError in the reference; here the native call reports it."""
    import pipeline_inputs as pi
    from longsom_b200.cli import splitbam
    bam, meta = pi.write_split_input("g2", str(tmp_path))
    with pytest.raises(RuntimeError, match="IndexError"):
        splitbam.main(["--bam", bam, "--meta", meta, "--id", "s", "--outdir", str(tmp_path), "--min_MQ", "0", "--n_trim", "500"])
    with pytest.raises(RuntimeError, match="cannot open"):
        splitbam.main(["--bam", bam + ".missing", "--meta", meta, "--id", "s", "--outdir", str(tmp_path)])


def test_split_streaming_rounds_and_flushes_do_not_change_the_output(tmp_path, built, monkeypatch):
    """The splitter streams: input in chunks of compressed bytes, outputs deflated whenever enough records are
    pending.  Tiny chunk sizes force hundreds of rounds, records that straddle rounds and mid-stream flushes; BAMs
    (record dumps), reports and region queries must not change."""
    import bai_query
    import pipeline_inputs as pi
    from longsom_b200.cli import splitbam
    bam, meta = pi.write_split_input("g1", str(tmp_path))
    dumps = {}
    for tag, env in (("big", {}), ("small", {"LS_SPLIT_READ_CHUNK": "70001", "LS_SPLIT_FLUSH_BYTES": "90000"}),
                     ("tiny", {"LS_SPLIT_READ_CHUNK": "4096", "LS_SPLIT_FLUSH_BYTES": "4096"})):
        out = os.path.join(str(tmp_path), tag)
        os.makedirs(out)
        for k in ("LS_SPLIT_READ_CHUNK", "LS_SPLIT_FLUSH_BYTES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        splitbam.main(["--bam", bam, "--meta", meta, "--id", "s", "--outdir", out, "--min_MQ", "60", "--n_trim", "5"])
        dumps[tag] = {fn: pi.dump_bam_records(os.path.join(out, fn)) for fn in sorted(os.listdir(out)) if fn.endswith(".bam")}
        if tag != "big":
            assert dumps[tag] == dumps["big"]
            fn = sorted(dumps[tag])[0]
            recs = [l.split("\t") for l in dumps[tag][fn][1:]]
            tid, pos = int(recs[len(recs) // 2][2]), int(recs[len(recs) // 2][3])
            assert bai_query.query(os.path.join(out, fn), tid, pos, pos + 2000) == \
                bai_query.query(os.path.join(str(tmp_path), "big", fn), tid, pos, pos + 2000)
    assert gold("g1", "split.s.Cancer.records.txt") == dumps["big"]["s.Cancer.bam"]
