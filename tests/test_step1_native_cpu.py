"""The native row passes of BaseCellCalling.step1 (csrc/host/ls_step1.cpp) against the Python restatement of the
reference's row loop (cli/step1.py:_parse_rows / _format_rows, itself pinned to the reference's golden outputs in
test_golden_cpu.py): byte-identical tables on the goldens and on randomised tables that reach every label of the
cascade; rows the reference would fail on end in the same Python exception (the native parser refuses them)."""
import gzip
import os

import numpy as np
import pytest

from longsom_b200 import bamio
from longsom_b200.cli import step1

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
A1, B1, A2, B2 = 0.21356677091082193, 104.95163748636298, 0.2474528917555431, 162.03696139428595


class ScipyEngine:
    last_stats = {}

    def betabinom_sf(self, k, n, a, b):
        from scipy.stats import betabinom
        k, n = np.asarray(k), np.asarray(n)
        return betabinom.sf(k - 0.001, n, a, b) if len(k) else np.zeros(0)


def _run(infile, outfile, fasta, native, **kw):
    old = os.environ.get("LONGSOM_STEP1_NATIVE")
    os.environ["LONGSOM_STEP1_NATIVE"] = "1" if native else "0"
    try:
        args = dict(min_ac_cells=2, min_ac_reads=3, min_cells=5, min_reads=5, min_cell_types=2, max_cell_types=1, fisher_cutoff=1)
        args.update(kw)
        return step1.variant_calling_step1(infile, outfile, fasta, A1, B1, A2, B2, args["min_ac_cells"], args["min_ac_reads"],
                                           args["min_cells"], args["min_reads"], args["min_cell_types"], args["max_cell_types"],
                                           args["fisher_cutoff"], ScipyEngine(), procs=1)
    finally:
        if old is None:
            del os.environ["LONGSOM_STEP1_NATIVE"]
        else:
            os.environ["LONGSOM_STEP1_NATIVE"] = old


def _used_native(monkeypatch):
    """Make the Python passes fail loudly, so that a silent fall-back cannot pass for the native path."""
    def boom(*a, **k):
        raise AssertionError("the Python passes ran")
    monkeypatch.setattr(step1, "_parse_rows", boom)


@pytest.mark.parametrize("case", ["g1", "g2", "g3"])
def test_native_equals_python_on_goldens(case, tmp_path, monkeypatch):
    merged = str(tmp_path / "merged.tsv")
    with gzip.open(os.path.join(GOLD, case, "merged.tsv.gz"), "rb") as f, open(merged, "wb") as o:
        o.write(f.read())
    # a reference with the contigs of the table, so that the context columns and both LC_ labels are exercised
    rng = np.random.default_rng(4)
    chroms = sorted({l.split("\t")[0] for l in open(merged) if not l.startswith("#")})
    top = max(int(l.split("\t")[1]) for l in open(merged) if not l.startswith("#"))
    fa = str(tmp_path / "ref.fa")
    seqs = ["".join(rng.choice(list("ACGTacgtN"), top - 3, p=[.2, .2, .2, .2, .04, .04, .04, .04, .04])).encode() for _ in chroms[:-1]]
    bamio.write_fasta(fa, chroms[:-1], seqs)   # the last contig is missing from the FASTA: '.' contexts
    py, nat = str(tmp_path / "py.tsv"), str(tmp_path / "nat.tsv")
    _run(merged, py, fa, native=False)
    _used_native(monkeypatch)
    _run(merged, nat, fa, native=True)
    assert open(nat, "rb").read() == open(py, "rb").read()


def _random_table(path, n_rows, seed, crlf=False):
    rng = np.random.default_rng(seed)
    types = ["Cancer", "Non-Cancer", "T_cell"]
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        f.write("##fileDate=x" + nl + "##INFO=whatever" + nl)
        f.write("\t".join(["#CHROM", "Start", "End", "REF", "INFO"] + types) + nl)
        pos = 0
        for i in range(n_rows):
            pos += int(rng.integers(1, 40))
            ref = str(rng.choice(["A", "C", "G", "T", "N", "A|C", "."], p=[.22, .22, .22, .22, .04, .04, .04]))
            cols = []
            for _t in types:
                u = rng.random()
                if u < 0.15:
                    cols.append("NA")
                    continue
                depth = int(rng.choice([0, 3, 8, 40, 400, 5000]))
                bc = np.zeros(6, np.int64)
                if depth:
                    main = "ACTG".find(ref) if ref in "ACTG" else int(rng.integers(0, 4))
                    bc[main] = depth
                    for _ in range(int(rng.integers(0, 4))):   # alternative alleles, indels
                        x = int(rng.integers(0, 6))
                        amount = int(rng.choice([1, 2, 3, 6, 30, depth // 3 + 1]))
                        amount = min(amount, int(bc[main]))
                        bc[main] -= amount
                        bc[x] += amount
                cc = np.minimum(bc, np.maximum((bc * rng.random(6)).astype(np.int64), (bc > 0).astype(np.int64)))
                dp, nc = int(bc.sum()), int(cc.sum())
                cols.append("|".join([str(dp), str(nc), ":".join(map(str, cc)), ":".join(map(str, bc)), ":".join(map(str, bc * 30)),
                                      ":".join(map(str, bc // 2)), ":".join(map(str, bc - bc // 2))]))
            f.write("\t".join(["chr%d" % (1 + i * 3 // n_rows), str(pos), str(pos), ref, "DP|NC|CC|BC|BQ|BCf|BCr"] + cols) + nl)
            if i == n_rows // 2:
                f.write("##a comment in the body" + nl)


@pytest.mark.parametrize("seed,crlf", [(1, False), (2, True), (3, False)])
def test_native_equals_python_on_random_tables(seed, crlf, tmp_path, monkeypatch):
    merged = str(tmp_path / "rand.tsv")
    _random_table(merged, 6000, seed, crlf)
    rng = np.random.default_rng(seed)
    fa = str(tmp_path / "ref.fa")
    # low-complexity reference: homopolymer contexts are common
    bamio.write_fasta(fa, ["chr1", "chr2"], ["".join(rng.choice(list("ACGT"), 200000, p=[.55, .15, .15, .15])).encode() for _ in range(2)])
    outs = []
    for kw in (dict(), dict(min_ac_cells=1, min_ac_reads=1, min_cell_types=1, max_cell_types=2, min_reads=3, min_cells=1),
               dict(min_ac_cells=100000), dict(min_ac_cells=1, min_ac_reads=100000)):
        py, nat = str(tmp_path / "py.tsv"), str(tmp_path / "nat.tsv")
        _run(merged, py, fa, native=False, **kw)
        with monkeypatch.context() as m:
            def boom(*a, **k):
                raise AssertionError("the Python passes ran")
            m.setattr(step1, "_parse_rows", boom)
            _run(merged, nat, fa, native=True, **kw)
        a, b = open(nat, "rb").read(), open(py, "rb").read()
        assert a == b
        outs.append(a)
    # the tables reach the labels of the cascade
    text = b"".join(outs).decode()
    for label in ("PASS", "Non-Significant", "Low-Significance", "Multi-allelic", "Low_cells", "Low_reads", "Multiple_cell_types",
                  "Min_cell_types", "Cell_type_noise", "Noisy_site", "LC_Upstream", "LC_Downstream"):
        assert label in text, label


def test_rows_the_reference_fails_on_fail_the_same_way(tmp_path):
    good = str(tmp_path / "good.tsv")
    _random_table(good, 200, 9)
    lines = open(good).read().split("\n")
    body0 = next(i for i, l in enumerate(lines) if l and not l.startswith("#"))
    cases = {
        "missing column": lambda l: "\t".join(l.split("\t")[:-1]),
        "non-numeric POS": lambda l: "\t".join([l.split("\t")[0], "12x"] + l.split("\t")[2:]),
        "six fields": lambda l: "\t".join(l.split("\t")[:5] + ["9|9|1:1:1:1:0:0|9:0:0:0:0:0|1:1|1:1"] + l.split("\t")[6:]),
        "empty line": lambda l: "",
    }
    for name, mutate in cases.items():
        bad = str(tmp_path / "bad.tsv")
        ls = list(lines)
        ls[body0 + 50] = mutate(ls[body0 + 50])
        open(bad, "w").write("\n".join(ls))
        errs = []
        for native in (False, True):
            with pytest.raises(Exception) as e:
                _run(bad, str(tmp_path / "o.tsv"), None, native=native)
            errs.append(type(e.value))
        assert errs[0] == errs[1], name


def test_native_waves_give_the_same_table(tmp_path, monkeypatch):
    """Small byte ranges (LONGSOM_STEP1_CHUNK_MB): many waves, each with its own K2 call, appended in order."""
    merged = str(tmp_path / "rand.tsv")
    _random_table(merged, 5000, 21)
    py, nat = str(tmp_path / "py.tsv"), str(tmp_path / "nat.tsv")
    _run(merged, py, None, native=False)
    monkeypatch.setenv("LONGSOM_STEP1_CHUNK_MB", "0.07")
    _used_native(monkeypatch)
    _run(merged, nat, None, native=True)
    assert open(nat, "rb").read() == open(py, "rb").read()
    assert not [f for f in os.listdir(str(tmp_path)) if f.endswith(".tmp")]
