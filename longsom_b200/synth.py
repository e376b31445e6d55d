"""Synthetic long-read scRNA data (test + bench infrastructure; not on the product path).

Thin ctypes wrapper over csrc/synth/ls_synth.c.  Shapes follow SURVEY.md 8(d); the five
BASELINE.json configs are available through `config(name, scale)`.
"""
import ctypes as C
import os

import numpy as np

from .batch import ReadBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libls_synth.so")


class _Params(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_contigs", C.c_int32), ("n_genes", C.c_int32), ("n_reads", C.c_int64),
        ("n_cells", C.c_int32), ("n_extra_cells", C.c_int32), ("frac_cancer", C.c_double),
        ("n_hot_genes", C.c_int32), ("hot_fraction", C.c_double), ("chrm_tid", C.c_int32),
        ("chrm_fraction", C.c_double), ("mean_len", C.c_double), ("sigma_len", C.c_double),
        ("p_mismatch", C.c_double), ("p_ins", C.c_double), ("p_del", C.c_double), ("p_softclip", C.c_double),
        ("p_no_cb", C.c_double), ("p_extra_cb", C.c_double), ("p_reverse", C.c_double), ("p_suppl", C.c_double),
        ("p_secondary", C.c_double), ("p_dup", C.c_double), ("p_qcfail", C.c_double), ("p_lowmapq", C.c_double),
        ("variants_per_gene", C.c_int32),
    ]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError("libls_synth.so not built; run __graft_entry__.build()")
        lib = C.CDLL(_LIB)
        lib.synth_plan_create.restype = C.c_void_p
        lib.synth_plan_create.argtypes = [C.POINTER(_Params), C.c_void_p]
        lib.synth_free.argtypes = [C.c_void_p]
        lib.synth_ref_len.restype = C.c_uint64
        lib.synth_ref_len.argtypes = [C.c_void_p]
        lib.synth_ref.restype = C.c_void_p
        lib.synth_ref.argtypes = [C.c_void_p]
        lib.synth_n_vars.restype = C.c_int32
        lib.synth_n_vars.argtypes = [C.c_void_p]
        lib.synth_n_exons.restype = C.c_int32
        lib.synth_n_exons.argtypes = [C.c_void_p]
        lib.synth_exons.argtypes = [C.c_void_p] * 4
        lib.synth_variants.argtypes = [C.c_void_p] * 7
        lib.synth_sizes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.synth_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 11
        lib.synth_headers.argtypes = [C.c_void_p] * 5
        _lib = lib
    return _lib


def barcode_of(cell_id, seed=0):
    """Deterministic 16-mer barcode of a dense cell id."""
    x = (int(cell_id) * 0x9E3779B97F4A7C15 + 0x1234567 + seed) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    x = (x * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 29
    s = []
    for _ in range(16):
        s.append("ACGT"[x & 3])
        x >>= 2
    # make collisions impossible: the id itself, base 4, in the last 10 letters
    t = int(cell_id)
    for i in range(10):
        s[15 - i] = "ACGT"[t & 3]
        t >>= 2
    return "".join(s)


class SynthData:
    """A generated data set: reference contigs + coordinate-sorted ReadBatch + truth tables."""

    def __init__(self, contig_names, contig_lens, ref, contig_off, batch, n_cells, n_extra_cells, frac_cancer,
                 exons, variants, params):
        self.contig_names = contig_names
        self.contig_lens = list(contig_lens)
        self.ref = ref                # uint8 array, all contigs back to back
        self.contig_off = contig_off  # int64 [n_contigs+1]
        self.batch = batch
        self.n_cells = n_cells
        self.n_extra_cells = n_extra_cells
        self.frac_cancer = frac_cancer
        self.exons = exons            # (tid, start, len)
        self.variants = variants      # dict of arrays
        self.params = params

    def contig_seq(self, t):
        return self.ref[int(self.contig_off[t]):int(self.contig_off[t + 1])]

    def contig_seqs(self):
        return {t: self.contig_seq(t) for t in range(len(self.contig_lens))}

    def n_cancer(self):
        return int(self.frac_cancer * self.n_cells)

    def cell_type(self, cell_id):
        return "Cancer" if cell_id < self.n_cancer() else "Non-Cancer"


class Plan:
    """Genome + gene models + sorted read order; reads are materialised on demand (all or a subset)."""

    def __init__(self, seed=1, contig_lens=(200000,), contig_names=None, n_genes=20, n_reads=2000, n_cells=100,
                 n_extra_cells=None, frac_cancer=0.4, n_hot_genes=0, hot_fraction=0.0, chrm=False,
                 chrm_fraction=0.02, mean_len=1500.0, sigma_len=0.35, p_mismatch=3e-3, p_ins=1e-3, p_del=1e-3,
                 p_softclip=0.3, p_no_cb=0.005, p_extra_cb=0.01, p_reverse=0.5, p_suppl=0.01, p_secondary=0.005,
                 p_dup=0.005, p_qcfail=0.002, p_lowmapq=0.07, variants_per_gene=3):
        lib = _load()
        self.contig_lens = [int(x) for x in contig_lens]
        nct = len(self.contig_lens)
        if contig_names is None:
            contig_names = ["chr%d" % (i + 1) for i in range(nct)]
            if chrm:
                contig_names[-1] = "chrM"
        self.contig_names = list(contig_names)
        if n_extra_cells is None:
            n_extra_cells = max(1, n_cells // 50)
        p = _Params()
        p.seed, p.n_contigs, p.n_genes, p.n_reads = seed, nct, n_genes, n_reads
        p.n_cells, p.n_extra_cells, p.frac_cancer = n_cells, n_extra_cells, frac_cancer
        p.n_hot_genes, p.hot_fraction = n_hot_genes, hot_fraction
        p.chrm_tid, p.chrm_fraction = (nct - 1 if chrm else -1), chrm_fraction
        p.mean_len, p.sigma_len = mean_len, sigma_len
        p.p_mismatch, p.p_ins, p.p_del, p.p_softclip = p_mismatch, p_ins, p_del, p_softclip
        p.p_no_cb, p.p_extra_cb, p.p_reverse, p.p_suppl = p_no_cb, p_extra_cb, p_reverse, p_suppl
        p.p_secondary, p.p_dup, p.p_qcfail, p.p_lowmapq = p_secondary, p_dup, p_qcfail, p_lowmapq
        p.variants_per_gene = variants_per_gene
        self.p = p
        self.n_reads, self.n_cells, self.n_extra_cells, self.frac_cancer = n_reads, n_cells, n_extra_cells, frac_cancer
        cl = np.array(self.contig_lens, np.int32)
        self._plan = lib.synth_plan_create(C.byref(p), cl.ctypes.data_as(C.c_void_p))
        if not self._plan:
            raise RuntimeError("synth_plan_create failed (too few genes?)")
        self.contig_off = np.zeros(nct + 1, np.int64)
        np.cumsum(cl.astype(np.int64), out=self.contig_off[1:])
        ref_len = lib.synth_ref_len(self._plan)
        self.ref = np.ctypeslib.as_array(C.cast(lib.synth_ref(self._plan), C.POINTER(C.c_uint8)),
                                         shape=(ref_len,)).copy()

    def close(self):
        if getattr(self, "_plan", None):
            _load().synth_free(self._plan)
            self._plan = None

    def __del__(self):
        self.close()

    def headers(self):
        """(tid, pos, gene_end, tlen) of every read in coordinate order, without materialising bases."""
        n = self.n_reads
        tid, pos, gend, tlen = (np.zeros(n, np.int32) for _ in range(4))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _load().synth_headers(self._plan, vp(tid), vp(pos), vp(gend), vp(tlen))
        return tid, pos, gend, tlen

    def materialize(self, sel=None):
        """ReadBatch of all reads (sel=None) or of the sorted-order indices in `sel` (ascending)."""
        lib = _load()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        if sel is not None:
            sel = np.ascontiguousarray(sel, np.int64)
            n = int(sel.shape[0])
            selp = vp(sel) if n else None
        else:
            n, selp = self.n_reads, None
        cigar_off = np.zeros(n + 1, np.uint32)
        base_off = np.zeros(n + 1, np.uint64)
        l_qseq = np.zeros(max(n, 1), np.int32)[:n]
        lib.synth_sizes(self._plan, selp, n, vp(cigar_off), vp(base_off), vp(l_qseq))
        n_cigar, n_bases = int(cigar_off[-1]), int(base_off[-1])
        tid, pos, cell = (np.zeros(max(n, 1), np.int32)[:n] for _ in range(3))
        flag = np.zeros(max(n, 1), np.uint16)[:n]
        mapq = np.zeros(max(n, 1), np.uint8)[:n]
        cigar = np.zeros(max(n_cigar, 1), np.uint32)[:n_cigar]
        seq4 = np.zeros(max(n_bases // 2, 1), np.uint8)[:n_bases // 2]
        qual = np.zeros(max(n_bases, 1), np.uint8)[:n_bases]
        uid = np.zeros(max(n, 1), np.int64)[:n]
        lib.synth_fill(self._plan, selp, n, vp(cigar_off), vp(base_off), vp(tid), vp(pos), vp(flag), vp(mapq), vp(cell),
                       vp(cigar), vp(seq4), vp(qual), vp(uid))
        batch = ReadBatch(tid, pos, flag, mapq, cell, cigar_off, cigar, base_off, l_qseq, seq4, qual)
        batch.uid = uid
        return batch

    def truth(self):
        lib = _load()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        ne = lib.synth_n_exons(self._plan)
        e_tid, e_start, e_len = np.zeros(ne, np.int32), np.zeros(ne, np.int32), np.zeros(ne, np.int32)
        lib.synth_exons(self._plan, vp(e_tid), vp(e_start), vp(e_len))
        nv = lib.synth_n_vars(self._plan)
        v = dict(tid=np.zeros(nv, np.int32), pos=np.zeros(nv, np.int32), alt=np.zeros(nv, np.uint8),
                 scope=np.zeros(nv, np.uint8), read_prob=np.zeros(nv, np.float32),
                 cell_frac=np.zeros(nv, np.float32))
        if nv:
            lib.synth_variants(self._plan, vp(v["tid"]), vp(v["pos"]), vp(v["alt"]), vp(v["scope"]), vp(v["read_prob"]),
                               vp(v["cell_frac"]))
        return (e_tid, e_start, e_len), v

    def data(self, sel=None):
        exons, variants = self.truth()
        return SynthData(self.contig_names, self.contig_lens, self.ref, self.contig_off, self.materialize(sel),
                         self.n_cells, self.n_extra_cells, self.frac_cancer, exons, variants,
                         dict(seed=int(self.p.seed), n_reads=self.n_reads))


def generate(**kw):
    """All reads of a plan as a SynthData."""
    plan = Plan(**kw)
    try:
        return plan.data()
    finally:
        plan.close()


# ---- BASELINE.json configs ---------------------------------------------------------------------
# Human-like relative contig lengths (chr1..chr22, X, Y) in Mb; scaled so the expressed span fits.
_HUMAN_MB = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50,
             156, 57]


def config(name, scale=1.0, seed=None):
    """Generator arguments of the five BASELINE.json configs; `scale` shrinks reads / genome / cells
    proportionally so that the same shape can run as a parity-sized test."""
    s = float(scale)
    if name == "C1":   # chr21 + chrM, 2e5 reads, 1k cells, 2 cell types
        return dict(seed=seed or 1, contig_lens=[max(60000, int(46_700_000 * s)), 16_600], contig_names=["chr21", "chrM"],
                    chrm=True, n_genes=max(8, int(1200 * s)), n_reads=max(200, int(200_000 * s)),
                    n_cells=max(20, int(1000 * min(1.0, s * 4))), frac_cancer=0.4)
    if name in ("C2", "C3"):  # whole transcriptome, 5e6 reads x 1.5 kb, 5k cells
        lens = [max(40000, int(mb * 1e6 * 0.06 * s)) for mb in _HUMAN_MB] + [16_600]
        return dict(seed=seed or 2, contig_lens=lens,
                    contig_names=["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY", "chrM"], chrm=True,
                    n_genes=max(30, int(12000 * s)), n_reads=max(500, int(5_000_000 * s)),
                    n_cells=max(50, int(5000 * min(1.0, s * 4))), frac_cancer=0.4)
    if name == "C4":   # hotspot stress: 20 genes at >1e5 reads/locus, 10k cells
        return dict(seed=seed or 4, contig_lens=[max(100000, int(30_000_000 * s))] * 4, n_genes=max(24, int(400 * s)),
                    n_hot_genes=20, hot_fraction=0.96, n_reads=max(2000, int(2_600_000 * s)),
                    n_cells=max(100, int(10000 * min(1.0, s * 4))), frac_cancer=0.4)
    if name == "C5":   # genotyping sweep base data set (sites / cells chosen by the caller)
        return dict(seed=seed or 5, contig_lens=[max(100000, int(40_000_000 * s))] * 6, n_genes=max(30, int(6000 * s)),
                    n_reads=max(1000, int(2_000_000 * s)), n_cells=max(100, int(20000 * min(1.0, s * 4))),
                    frac_cancer=0.4, variants_per_gene=8)
    raise KeyError(name)
