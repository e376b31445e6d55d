"""The native cursor merge (csrc/host/ls_merge.cpp) against the Python restatement of MergeBaseCellCounts.py:116-204
(cli/merge.py, pinned to the reference's goldens and quirks in test_golden_cpu.py): random sorted tables with gaps,
repeated and backwards positions, differing REF spellings (the frequency-ordered join), '.' references, a table without
rows, a blank line in the middle of a table; malformed rows end in the Python loop's exception."""
import ctypes as C
import os

import numpy as np
import pytest

from longsom_b200 import bamio
from longsom_b200.cli import merge

FMT = "DP|NC|CC|BC|BQ|BCf|BCr"
HDR = "##fileDate=x\n" + "\n".join("##INFO=%d" % i for i in range(7)) + "\n#CHROM\tStart\tREF\tINFO\tX\n"


def _tables(tmp_path, seed, n_tables=4, n_rows=3000):
    rng = np.random.default_rng(seed)
    chroms = ["chr1", "chr10", "chr2", "chrM", "chrX"]
    files = []
    for t in range(n_tables):
        fp = str(tmp_path / ("s.T%d.tsv" % t))
        with open(fp, "w") as f:
            f.write(HDR)
            if t == n_tables - 1 and seed % 2:
                files.append(fp)   # a table without rows
                continue
            for ch in chroms:
                if rng.random() < 0.2:
                    continue
                pos = np.sort(rng.integers(1, 4000, n_rows // len(chroms)))   # duplicates stay in: repeated positions
                for k, p in enumerate(pos.tolist()):
                    if rng.random() < 0.01:
                        p = max(1, p - int(rng.integers(1, 50)))   # a position that goes backwards
                    ref = str(rng.choice(["A", "a", "C", ".", "N"], p=[.5, .1, .2, .1, .1]))
                    f.write("%s\t%d\t%s\t%s\t%s\n" % (ch, p, ref, FMT, "t%d_%s_%d_%d" % (t, ch, p, k)))
                if t == 1 and ch == "chr2":
                    f.write("\nchr3\t1\tA\t%s\tafter_blank\n" % FMT)
                    break
        files.append(fp)
    return files


def _native(files, out):
    host = bamio._load_host()
    host.ls_merge_tables.restype = C.c_int
    host.ls_merge_tables.argtypes = [C.c_int32, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    paths = (C.c_char_p * len(files))(*[os.fsencode(p) for p in files])
    err = C.create_string_buffer(256)
    return host.ls_merge_tables(len(files), paths, os.fsencode(out), b"HEAD\n", 9, err, 256)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_native_merge_equals_python_loop(seed, tmp_path, monkeypatch):
    files = _tables(tmp_path, seed)
    nat, py = str(tmp_path / "nat.tsv"), str(tmp_path / "py.tsv")
    assert _native(files, nat) == 0
    monkeypatch.setenv("LONGSOM_MERGE_NATIVE", "0")
    merge.merge_cell_types_files(files, py)
    a = open(nat, "rb").read().split(b"\n", 1)[1]          # after "HEAD"
    b = open(py, "rb").read().split(b"\n", 9)[9]            # after the nine header lines
    assert a == b and len(a) > 1000
    monkeypatch.setenv("LONGSOM_MERGE_NATIVE", "1")
    merge.merge_cell_types_files(files, str(tmp_path / "cli.tsv"))
    assert open(str(tmp_path / "cli.tsv"), "rb").read().split(b"\n", 1)[1] == open(py, "rb").read().split(b"\n", 1)[1]


def test_malformed_rows_are_left_to_the_python_loop(tmp_path):
    files = _tables(tmp_path, 5, n_tables=2, n_rows=200)
    lines = open(files[0]).read().split("\n")
    for name, bad in (("four fields", "chr1\t5\tA\t" + FMT), ("six fields", "chr1\t5\tA\t%s\tx\ty" % FMT), ("text position", "chr1\tfive\tA\t%s\tx" % FMT)):
        ls = list(lines)
        ls[40] = bad
        open(files[0], "w").write("\n".join(ls))
        assert _native(files, str(tmp_path / "o.tsv")) == 1, name
        with pytest.raises(ValueError):
            merge.merge_cell_types_files(files, str(tmp_path / "o2.tsv"))


def test_site_list_reader_native_equals_python(tmp_path, monkeypatch):
    """step2's editing / PoN lists (BaseCellCalling.step2.py:197-221): the native reader and the Python loop give the same
    key set; a list the reference's loop fails on (and so empties) is empty either way."""
    from longsom_b200.cli import step2
    rng = np.random.default_rng(8)
    good = str(tmp_path / "sites.tsv")
    with open(good, "w", newline="") as f:
        f.write("#chrom\tpos\tinfo\n")
        for i in range(20000):
            ch = "chr%s" % rng.choice(["1", "2", "10", "X", "Un_gl000220"])
            pos = int(rng.integers(-5, 1 << 33)) if i % 97 == 0 else int(rng.integers(0, 250_000_000))
            end = "\r\n" if i % 3 == 0 else "\n"
            f.write(("%s\t%d\tx\ty" % (ch, pos) if i % 2 else "%s\t%d" % (ch, pos)) + end)
            if i == 5000:
                f.write("# a comment in the middle\n")

    def keys(path, native):
        monkeypatch.setenv("LONGSOM_STEP2_NATIVE", "1" if native else "0")
        ids = {}
        k = step2.site_list_keys(path, lambda names: [ids.setdefault(n, len(ids)) for n in names])
        inv = {v: n for n, v in ids.items()}
        return sorted((inv[int(x) >> 32], int(x) & 0xffffffff) for x in k.tolist())
    a, b = keys(good, True), keys(good, False)
    assert a == b and len(a) > 19000
    for name, bad_line in (("one column", "chr1\n"), ("text position", "chr1\tabc\n"), ("empty line", "\n")):
        bad = str(tmp_path / "bad.tsv")
        open(bad, "w").write(open(good).read() + bad_line)
        assert keys(bad, True) == [] and keys(bad, False) == [], name
    assert keys(str(tmp_path / "missing.tsv"), True) == [] and keys(str(tmp_path / "missing.tsv"), False) == []
