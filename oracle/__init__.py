"""CPU oracle of the LongSom SNV hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this package.  The product package longsom_b200/ never does.

Parity status: the reference has no golden vectors.  The oracle is pinned
against outputs of the reference scripts themselves, executed in the build container from
/root/reference over oracle/shims (see oracle/make_golden.py, tests/golden/).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError("oracle/liboracle.so not built; run __graft_entry__.build() or make -C oracle")
        lib = C.CDLL(_LIB)
        lib.oracle_pileup_count.restype = C.c_int64
        lib.oracle_pileup_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_genotype_count.restype = C.c_int64
        lib.oracle_genotype_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def pileup_count(batch, windows, params, threads=1):
    """BaseCellCounter.run_interval over every window.  Returns (SiteCounts, n_aligned)."""
    from longsom_b200.batch import SiteCounts  # container types only (no product compute)
    lib = _load()
    bs, ws, ps = batch.as_struct(), windows.as_struct(), params.as_struct()
    cap = int((windows.end.astype(np.int64) - windows.start.astype(np.int64)).sum()) if windows.n_windows else 0
    out = SiteCounts.empty(max(cap, 1))
    s = out.as_struct()
    nal = C.c_int64(0)
    n = lib.oracle_pileup_count(C.byref(bs), C.byref(ws), C.byref(ps), C.byref(s), int(threads), C.byref(nal))
    if n < 0 or n > cap:
        raise RuntimeError("oracle_pileup_count failed: %d" % n)
    return out.head(int(n)), int(nal.value)


def genotype_count(batch, site_tid, site_pos, alt_class, n_cells, min_bq=30, min_mq=255, max_depth=200000,
                   alt_only=False, bin_size=50000):
    """Pileup loop of SingleCellGenotype.run_interval; bins as build_dict_variants (floor(POS/bin), 1-based POS)."""
    from longsom_b200._lib import LsGenoParams
    lib = _load()
    site_tid = np.ascontiguousarray(site_tid, np.int32)
    site_pos = np.ascontiguousarray(site_pos, np.int32)
    alt_class = np.ascontiguousarray(alt_class, np.uint8)
    n = site_pos.shape[0]
    bin_id = (site_tid.astype(np.int64) << 32) | ((site_pos.astype(np.int64) + 1) // bin_size)
    bin_id = np.ascontiguousarray(bin_id, np.int64)
    dp = np.zeros((n, n_cells), np.int32)
    alt = np.zeros((n, n_cells), np.int32)
    gp = LsGenoParams(int(min_bq), int(min_mq), int(max_depth), 1 if alt_only else 0, int(bin_size), 0)
    bs = batch.as_struct()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.oracle_genotype_count(C.byref(bs), vp(site_tid), vp(site_pos), vp(alt_class), vp(bin_id), n, int(n_cells),
                              C.byref(gp), vp(dp), vp(alt))
    return dp, alt
