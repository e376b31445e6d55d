"""The native long-table writer of the genotype CLIs (csrc/host/ls_genorows.cpp) against the row logic of
SingleCellGenotype.py:128-214 / HCCVSingleCellGenotype.py:126-212 restated in Python (the same statements as the
LONGSOM_GENO_NATIVE=0 branch of cli/genotype.py): random touched pairs, both scripts, the chrM shortcut, p-values on both
sides of the cutoff, VAFs around 0.3."""
import ctypes as C

import numpy as np
import pytest

from longsom_b200 import bamio


def _python_rows(prefix, index, chrm, hits, cells, hccv, pvalue):
    out = []
    for s, pre in enumerate(prefix):
        for c, (bc, ctype) in enumerate(cells):
            DP, ALT, PV = hits[s].get(c, (0, 0, None))
            VAF, BETABIN, MUTATED = '.', '.', 'NoCoverage'
            if DP > 0:
                if not hccv:
                    VAF = round(ALT / DP, 4)
                if ALT > 0:
                    if hccv:
                        VAF = round(ALT / DP, 4)
                    if chrm[s]:
                        MUTATED = 'LowVAFChrM' if VAF < 0.3 else 'PASS'
                    else:
                        BETABIN = np.float64(PV)
                        MUTATED = 'PASS' if BETABIN < pvalue else 'BetaBin_problem'
                else:
                    if hccv:
                        VAF = float(0)
                    MUTATED = 'NoAltReads'
            group = [pre, bc, ctype, str(DP), str(ALT), str(VAF), str(BETABIN), str(MUTATED)]
            if not hccv:
                BIN = 1 if MUTATED == "PASS" else (3 if MUTATED == "NoCoverage" else 0)
                group += [str(BIN), index[s]]
            out.append('\t'.join(group) + '\n')
    return "".join(out).encode()


@pytest.mark.parametrize("hccv", [False, True])
def test_native_rows_equal_the_python_statements(hccv):
    host = bamio._load_host()
    host.ls_geno_rows.restype = C.c_int64
    host.ls_geno_rows.argtypes = [C.c_int32] + [C.c_void_p] * 9 + [C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    host.ls_geno_rows_free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(12 + hccv)
    n_sites, n_cells = 120, 300
    cells = [("BC%04d-1" % c, str(rng.choice(["Cancer", "Non-Cancer", "T cell"]))) for c in range(n_cells)]
    prefix = ["\t".join(["chr%s" % rng.choice(["1", "2", "M"]), str(p), str(p), "A", "G", "Cancer", str(int(rng.integers(1, 50)))])
              for p in rng.integers(1, 10 ** 8, n_sites)]
    index = [pre.split("\t")[0] + ":" + pre.split("\t")[1] + ":G" for pre in prefix]
    chrm = np.array([pre.startswith("chrM") and rng.random() < 0.7 for pre in prefix], np.uint8)
    hits, t_cell, t_dp, t_alt, t_p, lo, hi = [], [], [], [], [], [], []
    for s in range(n_sites):
        k = int(rng.integers(0, 40)) if s % 7 else 0
        cs = np.sort(rng.choice(n_cells, k, replace=False))
        h = {}
        lo.append(len(t_cell))
        for c in cs:
            dp = int(rng.choice([1, 2, 3, 7, 10, 33, 100, 1000]))
            alt = int(rng.choice([0, 0, 1, dp // 3, (3 * dp + 9) // 10, dp]))
            alt = min(alt, dp)
            p = float(np.round(rng.choice([0.0, 0.0001, 0.0099, 0.01, 0.0101, 0.05, 0.5, 1.0, rng.random()]), 4))
            h[int(c)] = (dp, alt, p)
            t_cell.append(int(c)); t_dp.append(dp); t_alt.append(alt); t_p.append(p)
        hi.append(len(t_cell))
        hits.append(h)
    want = _python_rows(prefix, index, chrm, hits, cells, hccv, 0.01)
    a = lambda x, dt: np.ascontiguousarray(np.array(x, dt))
    tc, td, ta, tp = a(t_cell, np.int32), a(t_dp, np.int32), a(t_alt, np.int32), a(t_p, np.float64)
    hlo, hhi = a(lo, np.int64), a(hi, np.int64)
    pre_a = (C.c_char_p * n_sites)(*[x.encode() for x in prefix])
    idx_a = None if hccv else (C.c_char_p * n_sites)(*[x.encode() for x in index])
    cell_a = (C.c_char_p * n_cells)(*[(b + "\t" + t).encode() for b, t in cells])
    text = C.c_void_p()
    n = host.ls_geno_rows(n_sites, pre_a, idx_a, chrm.ctypes.data, hlo.ctypes.data, hhi.ctypes.data, tc.ctypes.data, td.ctypes.data,
                          ta.ctypes.data, tp.ctypes.data, n_cells, cell_a, 1 if hccv else 0, 0.01, C.byref(text))
    assert n == len(want)
    got = C.string_at(text.value, n)
    host.ls_geno_rows_free(text)
    assert got == want
    for label in (b"NoCoverage", b"NoAltReads", b"PASS", b"BetaBin_problem", b"LowVAFChrM"):
        assert label in got
