"""Deterministic pipeline inputs (BAMs, FASTA, barcodes, editing / PoN / gnomAD tables) for the
golden-vector cases.  Used both by oracle/make_golden.py (which runs the REFERENCE scripts on
them, in the build container) and by the tests (which run the drop-in CLIs on the same files)."""
import gzip
import os

import numpy as np

from longsom_b200 import bamio, synth
from longsom_b200.batch import ReadBatch

CASES = {
    # name: synth arguments
    "g1": dict(seed=101, contig_lens=[120000, 16600], chrm=True, n_genes=8, n_reads=5000, n_cells=60,
               n_extra_cells=3, variants_per_gene=6, p_mismatch=5e-4, p_ins=2e-4, p_del=2e-4),
    "g2": dict(seed=202, contig_lens=[70000, 60000], n_genes=7, n_reads=2500, n_cells=40, n_extra_cells=2,
               variants_per_gene=5, p_ins=4e-3, p_del=4e-3),
    # 1 000 cells / 20 000 reads: the size SURVEY 8(d) asks for as a full reference-script run (core stages only, see
    # CORE_ONLY: the pure-Python pysam stand-in makes the reference scripts minutes per thousand reads)
    "g3": dict(seed=303, contig_lens=[260000], n_genes=14, n_reads=20000, n_cells=1000, n_extra_cells=5,
               variants_per_gene=6),
}
CORE_ONLY = ("g3",)  # BaseCellCounter x2, merge, step1, step2, SingleCellGenotype (All) only


def write_beds(case, workdir, d):
    """--bed / --bed_out files for BaseCellCounter (MakeWindows, BaseCellCounter.py:87-110): unsorted-free, touching
    and overlapping focus intervals (merge -d 1), one past the contig end (clipped by the intersect), and ignore
    intervals that split, trim and swallow focus pieces."""
    n0, l0 = d.contig_names[0], int(d.contig_lens[0])
    bed, bed_out = os.path.join(workdir, "focus.bed"), os.path.join(workdir, "ignore.bed")
    with open(bed, "w") as f:
        f.write("track name=focus\n")
        for s, e in ((0, 900), (900, 1800), (1801, 2500), (5000, 30000), (29000, 52000), (70000, 70001), (l0 - 700, l0 + 5000)):
            f.write("%s\t%d\t%d\n" % (n0, s, e))
        if len(d.contig_names) > 1:
            f.write("%s\t%d\t%d\n" % (d.contig_names[1], 10, min(9000, int(d.contig_lens[1]))))
    with open(bed_out, "w") as f:
        for s, e in ((1000, 1200), (4000, 6000), (20000, 20001), (51000, 60000), (l0 - 100, l0)):
            f.write("%s\t%d\t%d\n" % (n0, s, e))
    return bed, bed_out

# CB:Z suffix per case: "-1" (10x style; the HCCV script then never matches a barcode, quirk a12 of SURVEY.md)
# or none (HCCVSingleCellGenotype works)
CB_SUFFIX = {"g1": "-1", "g2": ""}

# alpha / beta the workflow passes (config/config.yaml:87-90 of the reference)
ALPHA1, BETA1 = 0.21356677091082193, 104.95163748636298
ALPHA2, BETA2 = 0.2474528917555431, 162.03696139428595


def _corner_reads(d):
    """Hand-built records for the CIGAR corner cases of SURVEY.md Appendix A, appended to contig 0:
    ref-skip followed by an insertion / deletion, padding, deletion followed by insertion,
    '=' / 'X' ops, IUPAC and '=' base codes, hard clips, qualities exactly at the thresholds."""
    M, I, D, N, S, H, P, EQ, X = range(9)
    base = 300  # positions near the start of contig 0 (no gene is placed before ~1)
    recs = []

    def add(pos, cig, cell, flag=0, mapq=60, qfill=40, codes=None, quals=None):
        ql = sum(l for op, l in cig if op in (M, I, S, EQ, X))
        c = np.array(codes if codes is not None else [1, 2, 4, 8] * (ql // 4 + 1), np.uint8)[:ql]
        q = np.array(quals if quals is not None else [qfill] * ql, np.uint8)[:ql]
        recs.append((pos, [(l << 4) | op for op, l in cig], cell, flag, mapq, c, q))
    for rep in range(8):  # enough depth for min_dp 5 / min_cc 5
        c = rep
        add(base, [(S, 3), (M, 20), (N, 40), (I, 2), (M, 20)], c)                 # N then I
        add(base, [(M, 20), (N, 40), (D, 3), (M, 20)], c, flag=16)                # N then D
        add(base + 2, [(M, 18), (P, 1), (I, 2), (M, 30)], c)                      # M P I
        add(base + 4, [(M, 10), (D, 4), (I, 3), (M, 30)], c, flag=16)             # D then I
        add(base + 5, [(H, 5), (EQ, 12), (X, 2), (EQ, 20), (H, 2)], c)            # = X with hard clips
        add(base + 6, [(M, 40)], c, codes=[0, 3, 5, 15, 1, 2, 4, 8] * 5)          # '=' / IUPAC / N codes
        add(base + 7, [(M, 30)], c, quals=[19, 20, 21, 29, 30, 31] * 5)           # thresholds of min_bq 20 / 30
        add(base + 8, [(M, 15), (D, 2), (D, 3), (M, 15)], c)                      # adjacent deletions
        add(base + 9, [(M, 25), (I, 4)], c)                                       # trailing insertion
        add(base + 9, [(I, 2), (M, 25)], c, flag=16)                              # leading insertion
    add(base + 1, [(M, 40)], -1)                       # no CB tag
    add(base + 1, [(M, 40)], 1, flag=0x800)            # supplementary
    add(base + 1, [(M, 40)], 2, flag=0x100)            # secondary
    add(base + 1, [(M, 40)], 3, flag=0x400)            # duplicate
    add(base + 1, [(M, 40)], 4, flag=0x200)            # QC fail
    add(base + 1, [(M, 40)], 5, flag=0x1)              # paired, not proper (orphan)
    add(base + 1, [(M, 40)], 6, flag=0x3)              # proper pair
    add(base + 1, [(M, 40)], 7, mapq=59)               # below min_mq 60
    return recs


def build_batch(case):
    """SynthData with the corner-case records merged in (coordinate order preserved)."""
    d = synth.generate(**CASES[case])
    if case != "g2":
        return d
    b = d.batch
    recs = _corner_reads(d)
    n0 = b.n_reads
    tid = np.concatenate([b.tid, np.zeros(len(recs), np.int32)])
    pos = np.concatenate([b.pos, np.array([r[0] for r in recs], np.int32)])
    cell = np.concatenate([b.cell, np.array([r[2] for r in recs], np.int32)])
    flag = np.concatenate([b.flag, np.array([r[3] for r in recs], np.uint16)])
    mapq = np.concatenate([b.mapq, np.array([r[4] for r in recs], np.uint8)])
    lq = np.concatenate([b.l_qseq, np.array([len(r[5]) for r in recs], np.int32)])
    cig_list = [b.cigar[b.cigar_off[i]:b.cigar_off[i + 1]] for i in range(n0)] + [np.array(r[1], np.uint32) for r in recs]
    codes = [None] * n0 + [r[5] for r in recs]
    quals = [None] * n0 + [r[6] for r in recs]
    order = np.lexsort((np.arange(len(pos)), pos, tid))
    cigar_off = np.zeros(len(pos) + 1, np.uint32)
    base_off = np.zeros(len(pos) + 1, np.uint64)
    for j, i in enumerate(order):
        cigar_off[j + 1] = cigar_off[j] + len(cig_list[i])
        base_off[j + 1] = base_off[j] + ((int(lq[i]) + 15) // 16) * 16
    cigar = np.zeros(int(cigar_off[-1]), np.uint32)
    seq4 = np.zeros(int(base_off[-1]) // 2, np.uint8)
    qual = np.zeros(int(base_off[-1]), np.uint8)
    for j, i in enumerate(order):
        cigar[cigar_off[j]:cigar_off[j + 1]] = cig_list[i]
        bo = int(base_off[j])
        if i < n0:
            so, l = int(b.base_off[i]), int(lq[i])
            qual[bo:bo + l] = b.qual[so:so + l]
            seq4[bo // 2:bo // 2 + (l + 1) // 2] = b.seq4[so // 2:so // 2 + (l + 1) // 2]
        else:
            c, q = codes[i], quals[i]
            qual[bo:bo + len(q)] = q
            for t, cd in enumerate(c):
                if t & 1:
                    seq4[(bo + t) >> 1] |= cd
                else:
                    seq4[(bo + t) >> 1] = cd << 4
    d.batch = ReadBatch(tid[order], pos[order], flag[order], mapq[order], cell[order], cigar_off, cigar, base_off,
                        lq[order], seq4, qual)
    return d


def cb_text(d, suffix="-1"):
    b = d.batch
    return lambda i: None if b.cell[i] < 0 else synth.barcode_of(int(b.cell[i])) + suffix


def write_inputs(case, workdir):
    """Creates <workdir>/{ref.fa(.fai), full.bam, bam/s.Cancer.bam, bam/s.Non-Cancer.bam, barcodes.tsv,
    editing.txt, editing.txt.gz, pon_SR.txt, pon_LR.txt, gnomad.tsv}; returns a dict of paths + the data."""
    os.makedirs(os.path.join(workdir, "bam"), exist_ok=True)
    d = build_batch(case)
    b = d.batch
    p = dict(ref=os.path.join(workdir, "ref.fa"), full=os.path.join(workdir, "full.bam"),
             cancer=os.path.join(workdir, "bam", "s.Cancer.bam"), normal=os.path.join(workdir, "bam", "s.Non-Cancer.bam"),
             meta=os.path.join(workdir, "barcodes.tsv"), editing=os.path.join(workdir, "editing.txt"),
             editing_gz=os.path.join(workdir, "editing.txt.gz"), pon_sr=os.path.join(workdir, "pon_SR.txt"),
             pon_lr=os.path.join(workdir, "pon_LR.txt"), gnomad=os.path.join(workdir, "gnomad.tsv"))
    bamio.write_fasta(p["ref"], d.contig_names, [d.contig_seq(t) for t in range(len(d.contig_lens))])
    suf = CB_SUFFIX.get(case, "-1")
    cb = cb_text(d, suf)
    bamio.write_bam(p["full"], d.contig_names, d.contig_lens, b, cb)
    ncan = d.n_cancer()
    # per-cell-type BAMs: reads of that type; the Cancer BAM also keeps CB-less and not-in-meta reads so
    # that BaseCellCounter's own gates are exercised (SplitBam would have dropped them)
    is_can = (b.cell < ncan)
    is_norm = (b.cell >= ncan) & (b.cell < d.n_cells)
    for path, sel in ((p["cancer"], np.nonzero(is_can | (b.cell >= d.n_cells))[0]), (p["normal"], np.nonzero(is_norm)[0])):
        sb = b.select(sel)
        sd_cb = lambda i, sb=sb: None if sb.cell[i] < 0 else synth.barcode_of(int(sb.cell[i])) + suf
        bamio.write_bam(path, d.contig_names, d.contig_lens, sb, sd_cb)
    with open(p["meta"], "w") as f:
        f.write("Index\tCell_type\n")
        for c in range(d.n_cells):
            f.write("%s%s\t%s\n" % (synth.barcode_of(c), suf, d.cell_type(c)))
    v = d.variants
    rng = np.random.default_rng(CASES[case]["seed"])
    ed = [(d.contig_names[t], int(q) + 1) for t, q, a, rp in zip(v["tid"], v["pos"], v["alt"], v["read_prob"]) if rp < 0.31 and rp != 0.5][::2]
    with open(p["editing"], "w") as f:
        f.write("#chrom\tpos\n")
        for c, q in ed:
            f.write("%s\t%d\tA\tG\n" % (c, q))
    with gzip.open(p["editing_gz"], "wt") as f:
        for c, q in ed:
            f.write("%s\t%d\tA\tG\n" % (c, q))
    pon = [(d.contig_names[t], int(q) + 1) for t, q in zip(v["tid"], v["pos"])]
    with open(p["pon_sr"], "w") as f:
        for c, q in pon[1::5]:
            f.write("%s\t%d\n" % (c, q))
        for _ in range(50):
            f.write("%s\t%d\n" % (d.contig_names[0], int(rng.integers(1, d.contig_lens[0]))))
    with open(p["pon_lr"], "w") as f:
        for c, q in pon[2::7]:
            f.write("%s\t%d\n" % (c, q))
    with open(p["gnomad"], "w") as f:
        f.write("#chrom\tpos\tref\talt\tAF\n")
        for k, (t, q, a) in enumerate(zip(v["tid"], v["pos"], v["alt"])):
            if k % 4 == 0:
                ref = chr(d.contig_seq(int(t))[int(q)]).upper()
                f.write("%s\t%d\t%s\t%s\t%g\n" % (d.contig_names[t], int(q) + 1, ref, chr(a), 0.2 if k % 8 == 0 else 0.001))
    return p, d


# ---- SplitBamCellTypes inputs / record dumps ----------------------------------------------
def write_split_input(case, workdir):
    """<workdir>/split_in.bam: the case's full BAM with nM / NH aux tags (some absent), three trailing reads
    without a reference, and <workdir>/split_meta.tsv whose cell types contain blanks and a repeated barcode.
    Returns (bam path, meta path)."""
    import struct
    d = build_batch(case)
    b = d.batch
    n = b.n_reads
    tid = b.tid.copy()
    pos = b.pos.copy()
    flag = b.flag.copy()
    tid[n - 3:] = -1          # unplaced reads sit at the end of a coordinate-sorted BAM
    pos[n - 3:] = -1
    flag[n - 3:] |= 4
    b2 = ReadBatch(tid, pos, flag, b.mapq, b.cell, b.cigar_off, b.cigar, b.base_off, b.l_qseq, b.seq4, b.qual)
    suf = CB_SUFFIX.get(case, "-1")

    def extra(i):
        out = b""
        if i % 11:
            out += b"nMC" + bytes([(i * 7919) % 9])
        if i % 13:
            out += b"NHC" + bytes([1 if (i * 31) % 10 else 1 + (i % 3)])
        if i % 5 == 0:
            out += b"xfi" + struct.pack("<i", i)
        return out
    bam = os.path.join(workdir, "split_in.bam")
    bamio.write_bam(bam, d.contig_names, d.contig_lens, b2, cb_text(d, suf), extra_tags=extra)
    meta = os.path.join(workdir, "split_meta.tsv")
    with open(meta, "w") as f:
        f.write("Index\tCell_type\n")
        for c in range(d.n_cells):
            ct = d.cell_type(c)
            if ct != "Cancer":
                ct = "T cell" if c % 3 == 0 else "Non-Cancer"
            f.write("%s%s\t%s\n" % (synth.barcode_of(c), suf, ct))
        f.write("%s%s\t%s\n" % (synth.barcode_of(0), suf, "T cell"))   # repeated barcode: the later row wins
    return bam, meta


def dump_bam_records(path):
    """One line per record: name, flag, tid, pos, mapq, qualities (phred+33) and the md5 of the raw record --
    enough to see what differs, strict enough to prove byte-for-byte passthrough.  Independent BAM walk
    (gzip module over the BGZF members), used on both the reference-side and the drop-in outputs."""
    import gzip
    import hashlib
    import struct
    with gzip.open(path, "rb") as f:
        buf = f.read()
    assert buf[:4] == b"BAM\1"
    l_text = struct.unpack_from("<I", buf, 4)[0]
    q = 8 + l_text
    n_ref = struct.unpack_from("<I", buf, q)[0]
    q += 4
    for _ in range(n_ref):
        q += 8 + struct.unpack_from("<I", buf, q)[0]
    lines = ["#header_md5=%s" % hashlib.md5(buf[:q]).hexdigest()]
    while q + 4 <= len(buf):
        bs = struct.unpack_from("<I", buf, q)[0]
        r0 = q + 4
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHI", buf, r0)
        qo = r0 + 32 + l_name + 4 * n_cig + (l_seq + 1) // 2
        qual = "".join(chr(min(x, 93) + 33) for x in buf[qo:qo + l_seq])
        lines.append("%s\t%d\t%d\t%d\t%d\t%s\t%s" % (buf[r0 + 32:r0 + 32 + l_name - 1].decode(), flag, tid, pos, mapq, qual,
                                                     hashlib.md5(buf[q:r0 + bs]).hexdigest()))
        q = r0 + bs
    return lines
