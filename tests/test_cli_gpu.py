"""GPU suite, end to end through the drop-in CLIs (the reference's own boundary: command line in,
TSV out).  Every script under workflow/scripts/ is run as a subprocess with the real CUDA engine
on regenerated inputs and must reproduce, byte for byte, what the UNMODIFIED reference scripts
wrote for the same inputs (tests/golden/, see oracle/make_golden.py).  ##fileDate is masked."""
import gzip
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
GOLD = os.path.join(ROOT, "tests", "golden")
SCRIPTS = os.path.join(ROOT, "workflow", "scripts")
CASES = [c for c in ("g1", "g2", "g3") if os.path.isdir(os.path.join(GOLD, c))]


def gold_lines(case, name, keep_date=False):
    if not os.path.exists(os.path.join(GOLD, case, name + ".gz")):
        return HashedGold(os.path.join(GOLD, case, name + ".sha256.json"))
    with gzip.open(os.path.join(GOLD, case, name + ".gz"), "rt") as f:
        return [l for l in f if keep_date or not l.startswith("##fileDate")]


def _align_cell_type_columns(got, want, header_row, first_col):
    """Bring the cell-type columns of `got` (from `first_col` on) into the golden's order; rows before `header_row`
    are comment lines and stay as they are."""
    hg, hw = got[header_row].rstrip("\n").split("\t"), want[header_row].rstrip("\n").split("\t")
    if hg == hw:
        return got
    assert sorted(hg[first_col:]) == sorted(hw[first_col:]) and hg[:first_col] == hw[:first_col], (hg, hw)
    perm = [hg.index(c) for c in hw]
    return got[:header_row] + ["\t".join(l.rstrip("\n").split("\t")[j] for j in perm) + "\n" for l in got[header_row:]]


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


def run_script(rel, args):
    cmd = [sys.executable, os.path.join(SCRIPTS, rel)] + [str(a) for a in args]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, "%s failed:\n%s" % (" ".join(cmd), r.stdout[-3000:])
    return r.stdout


@pytest.fixture(scope="module", params=CASES)
def pipeline(request, tmp_path_factory, built):
    """Runs the whole drop-in chain once per case (each stage consumes the previous stage's OWN output)."""
    import pipeline_inputs as pi
    case = request.param
    d = str(tmp_path_factory.mktemp("cli_" + case))
    p, data = pi.write_inputs(case, d)
    out = os.path.join(d, "out")
    os.makedirs(os.path.join(out, "counts"))
    os.makedirs(os.path.join(out, "counts_ac"))
    for name, bam in (("Cancer", p["cancer"]), ("Non-Cancer", p["normal"])):
        run_script("SNVCalling/BaseCellCounter.py", ["--bam", bam, "--ref", p["ref"], "--chrom", "all", "--out_folder",
                   os.path.join(out, "counts"), "--min_bq", 20, "--min_mq", 60, "--nprocs", 8, "--tmp_dir",
                   os.path.join(d, "tmp_" + name)])
        assert os.path.isdir(os.path.join(d, "tmp_" + name))  # Snakemake declares it as an output directory
    full = has_gold(case, "hccv.tsv")  # g3 holds the core stages only (oracle/make_golden.py, CORE_ONLY)
    if full:
        run_script("SNVCalling/BaseCellCounter.py", ["--bam", p["full"], "--ref", p["ref"], "--chrom", data.contig_names[0],
                   "--out_folder", os.path.join(out, "counts_ac"), "--id", "full.ac", "--min_bq", 30, "--min_mq", 0, "--min_ac", 2,
                   "--min_dp", 3, "--min_cc", 2, "--bin", 30000, "--tmp_dir", os.path.join(d, "tmp_ac")])
    run_script("SNVCalling/MergeBaseCellCounts.py", ["--tsv_folder", os.path.join(out, "counts"), "--outfile",
               os.path.join(out, "merged.tsv")])
    # The merged table's cell-type column order is the glob order of the per-cell-type tables, which is filesystem
    # dependent in the reference too (quirk Q10).  The goldens were written under one order: present step1 with that
    # order (a pure column permutation of our own merge output), so that every later table is comparable byte for
    # byte and no comparison ever has to be skipped.  test_merge checks the merge's own output.
    merged = os.path.join(out, "merged.tsv")
    os.replace(merged, os.path.join(out, "merged.own_order.tsv"))
    with open(os.path.join(out, "merged.own_order.tsv")) as f:
        lines = f.readlines()
    lines = _align_cell_type_columns(lines, gold_lines(case, "merged.tsv", keep_date=True), 8, 5)
    with open(merged, "w") as f:
        f.writelines(lines)
    run_script("SNVCalling/BaseCellCalling.step1.py", ["--infile", os.path.join(out, "merged.tsv"), "--outfile",
               os.path.join(out, "s"), "--ref", p["ref"], "--min_cell_types", 2, "--min_ac_reads", 3, "--min_ac_cells", 2,
               "--alpha1", pi.ALPHA1, "--beta1", pi.BETA1, "--alpha2", pi.ALPHA2, "--beta2", pi.BETA2])
    run_script("SNVCalling/BaseCellCalling.step2.py", ["--infile", os.path.join(out, "s.calling.step1.tsv"), "--outfile",
               os.path.join(out, "s"), "--editing", p["editing"], "--pon_SR", p["pon_sr"], "--pon_LR", p["pon_lr"],
               "--gnomAD_db", p["gnomad"], "--gnomAD_max", 0.01, "--min_distance", 0])
    if full:
        run_script("SNVCalling/BaseCellCalling.step2.py", ["--infile", os.path.join(out, "s.calling.step1.tsv"), "--outfile",
                   os.path.join(out, "s.gz"), "--editing", p["editing_gz"], "--pon_SR", p["pon_sr"], "--pon_LR", "--gnomAD_db",
                   p["gnomad"], "--gnomAD_max", 0.01, "--min_distance", 5])
    cand = os.path.join(out, "candidates.tsv")
    with gzip.open(os.path.join(GOLD, case, "candidates.tsv.gz"), "rb") as f, open(cand, "wb") as o:
        o.write(f.read())
    for flag in (("All", "Alt") if full else ("All",)):
        run_script("CellClustering/SingleCellGenotype.py", ["--bam", p["full"], "--infile", cand, "--ref", p["ref"], "--meta",
                   p["meta"], "--fusions", "--outfile", os.path.join(out, "geno_" + flag), "--alt_flag", flag, "--nprocs", 4,
                   "--min_mq", 60, "--pvalue", 0.01, "--alpha2", pi.ALPHA2, "--beta2", pi.BETA2, "--chrM_contaminant", "True",
                   "--tmp_dir", os.path.join(d, "tmp_g" + flag)])
    if full:
        run_script("CellTypeReannotation/HCCVSingleCellGenotype.py", ["--bam", p["full"], "--infile", cand, "--ref", p["ref"],
                   "--meta", p["meta"], "--outfile", os.path.join(out, "hccv.tsv"), "--alt_flag", "All", "--nprocs", 4, "--min_mq",
                   60, "--pvalue", 0.01, "--chrM_contaminant", "True", "--tmp_dir", os.path.join(d, "tmp_h")])
    return case, out


def test_base_cell_counter(pipeline):
    case, out = pipeline
    for name in ("Cancer", "Non-Cancer"):
        assert_same(file_lines(os.path.join(out, "counts", "s.%s.tsv" % name)), gold_lines(case, "counts.%s.tsv" % name),
                    "BaseCellCounter " + name)
    if has_gold(case, "counts.full_ac.tsv"):
        assert_same(file_lines(os.path.join(out, "counts_ac", "full.ac.tsv")), gold_lines(case, "counts.full_ac.tsv"),
                    "BaseCellCounter --min_ac 2 --bin 30000")


def test_base_cell_counter_bed(pipeline, tmp_path):
    """--bed / --bed_out through the drop-in CLI (MakeWindows, BaseCellCounter.py:87-110)."""
    import pipeline_inputs as pi
    case, out = pipeline
    if not has_gold(case, "counts.full_bed.tsv"):
        return  # the --bed goldens were generated for g1
    d = str(tmp_path)
    p, data = pi.write_inputs(case, d)
    bed, bed_out = pi.write_beds(case, d, data)
    for name, extra in (("bed", ["--bed", bed, "--chrom", "all"]),
                        ("bed_out", ["--bed", bed, "--bed_out", bed_out, "--chrom", data.contig_names[0]]),
                        ("bedout_only", ["--bed_out", bed_out, "--bin", 20000, "--chrom", "all"])):
        run_script("SNVCalling/BaseCellCounter.py", ["--bam", p["full"], "--ref", p["ref"], "--out_folder", d, "--id",
                   "full." + name, "--min_bq", 20, "--min_mq", 60, "--min_dp", 3, "--min_cc", 2, "--tmp_dir",
                   os.path.join(d, "tmp_" + name)] + extra)
        assert_same(file_lines(os.path.join(d, "full.%s.tsv" % name)), gold_lines(case, "counts.full_%s.tsv" % name),
                    "BaseCellCounter --" + name)


def test_merge(pipeline):
    case, out = pipeline
    got, want = file_lines(os.path.join(out, "merged.own_order.tsv")), gold_lines(case, "merged.tsv")
    assert_same(_align_cell_type_columns(got, want, 7, 5), want, "MergeBaseCellCounts")


def test_step1(pipeline):
    case, out = pipeline
    got, want = file_lines(os.path.join(out, "s.calling.step1.tsv")), gold_lines(case, "step1.tsv")
    assert_same(got, want, "BaseCellCalling.step1")


def test_step2(pipeline):
    case, out = pipeline
    got, want = file_lines(os.path.join(out, "s.calling.step2.tsv")), gold_lines(case, "step2.tsv")
    assert_same(got, want, "BaseCellCalling.step2")
    if has_gold(case, "step2_gz.tsv"):
        assert_same(file_lines(os.path.join(out, "s.gz.calling.step2.tsv")), gold_lines(case, "step2_gz.tsv"),
                    "BaseCellCalling.step2 with a gz editing list (filter silently off, Q9)")


def test_single_cell_genotype(pipeline):
    case, out = pipeline
    for flag in ("All", "Alt"):
        if not has_gold(case, "geno_%s.DpMatrix.tsv" % flag):
            continue
        for suf in ("SingleCellGenotype", "DpMatrix", "AltMatrix", "VAFMatrix", "BinaryMatrix"):
            assert_same(file_lines(os.path.join(out, "geno_%s.%s.tsv" % (flag, suf))),
                        gold_lines(case, "geno_%s.%s.tsv" % (flag, suf)), "SingleCellGenotype %s %s" % (flag, suf))


def test_hccv_single_cell_genotype(pipeline):
    case, out = pipeline
    if has_gold(case, "hccv.tsv"):
        assert_same(file_lines(os.path.join(out, "hccv.tsv")), gold_lines(case, "hccv.tsv"), "HCCVSingleCellGenotype")


def test_multi_device_env_gives_identical_table(pipeline, tmp_path):
    """LONGSOM_GPUS shards windows over engines; with one physical GPU listed twice the shards run on the
    same device and the concatenated table must be identical to the single-engine one."""
    import pipeline_inputs as pi
    case, out = pipeline
    d = str(tmp_path)
    p, data = pi.write_inputs(case, d)
    env = dict(os.environ, LONGSOM_GPUS="0,0,0")
    cmd = [sys.executable, os.path.join(SCRIPTS, "SNVCalling/BaseCellCounter.py"), "--bam", p["cancer"], "--ref", p["ref"],
           "--chrom", "all", "--out_folder", d, "--min_bq", "20", "--min_mq", "60", "--tmp_dir", os.path.join(d, "tmp")]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert_same(file_lines(os.path.join(d, "s.Cancer.tsv")), gold_lines(case, "counts.Cancer.tsv"), "3-shard BaseCellCounter")
