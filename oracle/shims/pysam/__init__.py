"""pysam compatibility shim -- TEST INFRASTRUCTURE (oracle), never imported by the product.

pysam / htslib are not installed in the build container and cannot be (no network).  To run the
UNMODIFIED reference scripts from /root/reference and obtain golden outputs, this module
restates, in pure Python, exactly the slice of the pysam API those scripts touch
 with the htslib pileup semantics of SURVEY.md Appendix A:

  AlignmentFile(path).pileup(contig, start, stop, min_base_quality=, min_mapping_quality=,
                             ignore_overlaps=, max_depth=)            -> iterator of PileupColumn
  PileupColumn.pos / .get_num_aligned() / .get_query_names() / .get_query_qualities() /
               .get_query_sequences(mark_matches=, add_indels=) / .pileups
  PileupRead.alignment.{opt, is_secondary, is_duplicate, is_supplementary, is_reverse}
  FastaFile(path).{references, get_reference_length, fetch, close}

The engine below is a literal re-statement of htslib's bam_plp_push / bam_plp_next /
resolve_cigar2 state machine (column by column, per-read CIGAR cursor), deliberately
structured differently from both the CUDA kernels (tile/segment events) and the C oracle
(per-window event lists), so that agreement between the three is meaningful.
It has its own BAM decoder (gzip module over the BGZF members + struct).
"""
import gzip
import struct

_NT16 = "=ACMGRSVTWYHKDBN"
BAM_CMATCH, BAM_CINS, BAM_CDEL, BAM_CREF_SKIP, BAM_CSOFT_CLIP, BAM_CHARD_CLIP, BAM_CPAD, BAM_CEQUAL, BAM_CDIFF = range(9)
_REF_OPS = (BAM_CMATCH, BAM_CDEL, BAM_CREF_SKIP, BAM_CEQUAL, BAM_CDIFF)
_MATCH_OPS = (BAM_CMATCH, BAM_CEQUAL, BAM_CDIFF)

BAM_FPAIRED, BAM_FPROPER_PAIR, BAM_FUNMAP, BAM_FREVERSE = 0x1, 0x2, 0x4, 0x10
BAM_FSECONDARY, BAM_FQCFAIL, BAM_FDUP, BAM_FSUPPLEMENTARY = 0x100, 0x200, 0x400, 0x800


class AlignedSegment:
    __slots__ = ("tid", "pos", "mapq", "flag", "name", "cigar", "seq", "qual", "tags", "end", "raw", "qoff")

    @property
    def cigartuples(self):
        return list(self.cigar) if self.cigar else None

    @property
    def query_qualities(self):
        """array('B') copy of the base qualities, None when absent (0xff filled), like pysam."""
        import array
        if len(self.qual) == 0 or self.qual[0] == 0xff:
            return None
        return array.array('B', self.qual)

    @query_qualities.setter
    def query_qualities(self, q):
        q = bytes(bytearray(q))
        if len(q) != len(self.qual):
            raise ValueError("quality and sequence mismatch: %i != %i" % (len(q), len(self.qual)))
        self.qual = q

    def to_record(self):
        """block_size + record bytes as bam_write1 emits them (the stored record with the current qualities)."""
        return self.raw[:self.qoff] + bytes(self.qual) + self.raw[self.qoff + len(self.qual):]

    def opt(self, tag):
        return self.tags[tag]  # KeyError when absent, like pysam

    def has_tag(self, tag):
        return tag in self.tags

    @property
    def is_secondary(self):
        return bool(self.flag & BAM_FSECONDARY)

    @property
    def is_duplicate(self):
        return bool(self.flag & BAM_FDUP)

    @property
    def is_supplementary(self):
        return bool(self.flag & BAM_FSUPPLEMENTARY)

    @property
    def is_reverse(self):
        return bool(self.flag & BAM_FREVERSE)

    @property
    def query_name(self):
        return self.name

    @property
    def reference_start(self):
        return self.pos

    @property
    def mapping_quality(self):
        return self.mapq


def _parse_aux(buf, p, end):
    tags = {}
    while p + 3 <= end:
        tag = buf[p:p + 2].decode()
        ty = chr(buf[p + 2])
        p += 3
        if ty == "A":
            tags[tag] = chr(buf[p]); p += 1
        elif ty in "cC":
            tags[tag] = struct.unpack_from("<b" if ty == "c" else "<B", buf, p)[0]; p += 1
        elif ty in "sS":
            tags[tag] = struct.unpack_from("<h" if ty == "s" else "<H", buf, p)[0]; p += 2
        elif ty in "iI":
            tags[tag] = struct.unpack_from("<i" if ty == "i" else "<I", buf, p)[0]; p += 4
        elif ty == "f":
            tags[tag] = struct.unpack_from("<f", buf, p)[0]; p += 4
        elif ty in "ZH":
            e = buf.index(b"\0", p)
            tags[tag] = buf[p:e].decode(); p = e + 1
        elif ty == "B":
            sub = chr(buf[p]); n = struct.unpack_from("<I", buf, p + 1)[0]
            size = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
            p += 5 + size * n
        else:
            break
    return tags


class AlignmentFile:
    """Whole-file in-memory BAM reader (fine for the small golden-vector inputs)."""
    _cache = {}

    def __init__(self, path, mode="rb", template=None, **kw):
        self.filename = path
        self._mode = mode
        if mode == "wb":  # SplitBamCellTypes.py:56 -- records are buffered and written on close()
            self._header = template._header
            self.references, self.lengths = template.references, template.lengths
            self._out = []
            return
        if path not in AlignmentFile._cache:
            AlignmentFile._cache[path] = self._load(path)
        self.references, self.lengths, self._reads, self._header = AlignmentFile._cache[path]

    def write(self, read):
        self._out.append(read.to_record())

    def fetch(self, *a, **kw):
        """fetch() without a region on an indexed file: every record placed on a reference, in file order."""
        if a or kw:
            raise NotImplementedError("shim: only fetch() without arguments")
        import copy
        for r in self._reads:
            if r.tid >= 0:
                yield copy.copy(r)  # the caller may edit qualities; the cached record must stay pristine

    @staticmethod
    def _load(path):
        with gzip.open(path, "rb") as f:
            buf = f.read()
        if buf[:4] != b"BAM\1":
            raise ValueError("not a BAM file: %s" % path)
        l_text = struct.unpack_from("<I", buf, 4)[0]
        p = 8 + l_text
        n_ref = struct.unpack_from("<I", buf, p)[0]
        p += 4
        names, lens = [], []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<I", buf, p)[0]
            names.append(buf[p + 4:p + 4 + l_name - 1].decode())
            lens.append(struct.unpack_from("<I", buf, p + 4 + l_name)[0])
            p += 8 + l_name
        reads = []
        n = len(buf)
        hdr_end = p
        while p + 4 <= n:
            bs = struct.unpack_from("<I", buf, p)[0]
            r0 = p + 4
            tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, _nt, _np, _tl = struct.unpack_from("<iiBBHHHIiii", buf, r0)
            q = r0 + 32
            a = AlignedSegment()
            a.tid, a.pos, a.mapq, a.flag = tid, pos, mapq, flag
            a.name = buf[q:q + l_name - 1].decode()
            q += l_name
            a.cigar = [(c & 15, c >> 4) for c in struct.unpack_from("<%dI" % n_cig, buf, q)]
            q += 4 * n_cig
            sb = buf[q:q + (l_seq + 1) // 2]
            q += (l_seq + 1) // 2
            a.seq = bytes((sb[i >> 1] >> 4) if not (i & 1) else (sb[i >> 1] & 15) for i in range(l_seq))
            a.qual = buf[q:q + l_seq]
            a.raw = buf[p:r0 + bs]
            a.qoff = q - p
            q += l_seq
            a.tags = _parse_aux(buf, q, r0 + bs)
            rlen = sum(l for op, l in a.cigar if op in _REF_OPS)
            a.end = pos + rlen  # raw rlen, as bam_plp_push uses
            reads.append(a)
            p = r0 + bs
        return names, lens, reads, bytes(buf[:hdr_end])

    def close(self):
        if self._mode != "wb" or self._out is None:
            return
        import zlib
        data = self._header + b"".join(self._out)
        with open(self.filename, "wb") as f:
            for lo in range(0, len(data), 0xff00):
                chunk = data[lo:lo + 0xff00]
                co = zlib.compressobj(6, zlib.DEFLATED, -15)
                body = co.compress(chunk) + co.flush()
                f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25))
                f.write(body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
            f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
        self._out = None

    def get_tid(self, name):
        return self.references.index(name)

    def fetch_region(self, tid, start, stop):
        """Records overlapping [start, stop) in file order (what the BAI query yields)."""
        for a in self._reads:
            if a.tid != tid:
                continue
            if a.pos >= stop:
                break
            endpos = a.end if a.end > a.pos else a.pos + 1  # bam_endpos
            if endpos > start:
                yield a

    def pileup(self, contig=None, start=None, stop=None, min_base_quality=13, min_mapping_quality=0,
               ignore_overlaps=True, max_depth=8000, flag_filter=BAM_FUNMAP | BAM_FSECONDARY | BAM_FQCFAIL | BAM_FDUP,
               ignore_orphans=True, truncate=False, stepper="samtools", **kw):
        tid = self.get_tid(contig)
        if start is None:
            start = 0
        if start < 0:
            raise ValueError("start out of range (%i)" % start)
        if stop is None:
            stop = self.lengths[tid]

        def source():  # pysam __advance_samtools
            for a in self.fetch_region(tid, start, stop):
                if a.flag & flag_filter:
                    continue
                if a.mapq < min_mapping_quality:
                    continue
                if ignore_orphans and (a.flag & BAM_FPAIRED) and not (a.flag & BAM_FPROPER_PAIR):
                    continue
                yield a
        return _PileupEngine(source(), max_depth, min_base_quality)


Samfile = AlignmentFile


def index(path, *args):
    """pysam.index(): the golden pipeline only needs the side effect (a .bai next to the BAM); the shim readers
    scan whole files, so a header-only index is written."""
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", 0))


class _Node:
    __slots__ = ("b", "beg", "end", "k", "x", "y")


class PileupRead:
    __slots__ = ("alignment", "query_position_or_next", "is_del", "is_refskip", "indel")

    @property
    def query_position(self):
        return None if self.is_del else self.query_position_or_next


class PileupColumn:
    def __init__(self, tid, pos, entries, min_bq):
        self.reference_id = tid
        self.pos = self.reference_pos = pos
        self._entries = entries  # list of PileupRead (all, before the base-quality skip)
        self._min_bq = min_bq

    def _visible(self):
        # pysam pileup_base_qual_skip: quality of qpos (next base for D/N), 0 if past the sequence
        out = []
        for p in self._entries:
            b = p.alignment
            q = b.qual[p.query_position_or_next] if p.query_position_or_next < len(b.qual) else 0
            if q < self._min_bq:
                continue
            out.append((p, q))
        return out

    def get_num_aligned(self):
        return len(self._visible())

    @property
    def nsegments(self):
        return len(self._entries)

    def get_query_names(self):
        return [p.alignment.name for p, _ in self._visible()]

    def get_query_qualities(self):
        return [q for _, q in self._visible()]

    @property
    def pileups(self):
        return [p for p, _ in self._visible()]

    def get_query_sequences(self, mark_matches=False, mark_ends=False, add_indels=False):
        res = []
        for p, _ in self._visible():
            b = p.alignment
            rev = b.is_reverse
            s = ""
            if not p.is_del:
                qp = p.query_position_or_next
                cc = _NT16[b.seq[qp]] if qp < len(b.seq) else "N"
                # mark_matches needs a reference sequence, which the scripts never supply
                if cc == "=":
                    cc = "," if rev else "."
                elif rev:
                    cc = cc.lower()
                s += cc
            elif add_indels:
                if p.is_refskip:
                    s += "<" if rev else ">"
                else:
                    s += "*"
            if add_indels:
                if p.indel > 0:
                    s += "+%d" % p.indel
                    for j in range(1, p.indel + 1):
                        qj = p.query_position_or_next + j
                        cc = _NT16[b.seq[qj]] if qj < len(b.seq) else "N"
                        s += cc.lower() if rev else cc
                elif p.indel < 0:
                    s += "-%d" % (-p.indel)
                    s += ("n" if rev else "N") * (-p.indel)
            res.append(s)
        return res


class _PileupEngine:
    """htslib bam_plp_t: push records, emit columns (bam_plp_push / bam_plp64_next / bam_plp64_auto)."""

    def __init__(self, source, maxcnt, min_bq):
        self.src = source
        self.maxcnt = maxcnt
        self.min_bq = min_bq
        self.nodes = []        # live nodes, file order (the linked list head..tail, tail excluded)
        self.cnt = 1           # mempool count: the empty tail node is always allocated
        self.tid, self.pos = 0, 0
        self.max_tid, self.max_pos = -1, -1
        self.is_eof = False

    def __iter__(self):
        return self

    # -- bam_plp_push -----------------------------------------------------------------------
    def _push(self, b):
        if b is None:
            self.is_eof = True
            return
        if b.tid < 0 or (b.flag & BAM_FUNMAP):
            return
        if self.maxcnt and self.tid == b.tid and self.pos == b.pos and self.cnt > self.maxcnt:
            return  # depth cap: drop
        n = _Node()
        n.b, n.beg, n.end = b, b.pos, b.end
        n.k = -1
        n.x = n.y = 0
        self.max_tid, self.max_pos = b.tid, n.beg
        if n.end > self.pos or b.tid > self.tid:
            self.nodes.append(n)
            self.cnt += 1

    # -- resolve_cigar2 -----------------------------------------------------------------------
    @staticmethod
    def _resolve(n, pos):
        b = n.b
        cig = b.cigar
        ncig = len(cig)
        if n.k == -1:
            if ncig == 1:
                if cig[0][0] in _MATCH_OPS:
                    n.k, n.x, n.y = 0, b.pos, 0
            else:
                n.x, n.y = b.pos, 0
                k = 0
                while k < ncig:
                    op, l = cig[k]
                    if op in _REF_OPS:
                        break
                    if op in (BAM_CINS, BAM_CSOFT_CLIP):
                        n.y += l
                    k += 1
                n.k = k
        else:
            op, l = cig[n.k]
            if pos - n.x >= l:
                op2 = cig[n.k + 1][0]
                if op in _MATCH_OPS:
                    n.y += l
                n.x += l
                if op2 in _REF_OPS:
                    n.k += 1
                else:
                    k = n.k + 1
                    while k < ncig:
                        o, ll = cig[k]
                        if o in _REF_OPS:
                            break
                        if o in (BAM_CINS, BAM_CSOFT_CLIP):
                            n.y += ll
                        k += 1
                    n.k = k
        op, l = cig[n.k]
        p = PileupRead()
        p.alignment = b
        p.is_del = p.is_refskip = False
        p.indel = 0
        if n.x + l - 1 == pos and n.k + 1 < ncig:
            op2, l2 = cig[n.k + 1]
            if op2 == BAM_CDEL and op != BAM_CDEL:
                p.indel = -l2
                for k in range(n.k + 2, ncig):
                    o, ll = cig[k]
                    if o == BAM_CDEL:
                        p.indel -= ll
                    else:
                        break
            elif op2 == BAM_CINS:
                p.indel = l2
                for k in range(n.k + 2, ncig):
                    o, ll = cig[k]
                    if o == BAM_CINS:
                        p.indel += ll
                    elif o != BAM_CPAD:
                        break
            elif op2 == BAM_CPAD and n.k + 2 < ncig:
                l3 = 0
                for k in range(n.k + 2, ncig):
                    o, ll = cig[k]
                    if o == BAM_CINS:
                        l3 += ll
                    elif o in _REF_OPS:
                        break
                if l3 > 0:
                    p.indel = l3
        if op in _MATCH_OPS:
            p.query_position_or_next = n.y + (pos - n.x)
        else:
            p.is_del = True
            p.query_position_or_next = n.y
            p.is_refskip = op == BAM_CREF_SKIP
        return p

    # -- bam_plp64_next -----------------------------------------------------------------------
    def _next(self):
        if self.is_eof and not self.nodes:
            return None
        while self.is_eof or self.max_tid > self.tid or (self.max_tid == self.tid and self.max_pos > self.pos):
            entries = []
            keep = []
            for n in self.nodes:
                if n.b.tid < self.tid or (n.b.tid == self.tid and n.end <= self.pos):
                    self.cnt -= 1  # mp_free
                    continue
                if n.b.tid == self.tid and n.beg <= self.pos:
                    entries.append(self._resolve(n, self.pos))
                keep.append(n)
            self.nodes = keep
            col = (self.tid, self.pos, entries)
            if self.nodes:
                head = self.nodes[0]
                if self.tid < head.b.tid:
                    self.tid, self.pos = head.b.tid, head.beg
                elif self.pos < head.beg:
                    self.pos = head.beg
                else:
                    self.pos += 1
            else:
                # htslib reads the (stale) tail node here; with an empty list the next push decides.
                self.pos += 1
            if entries:
                return col
            if self.is_eof and not self.nodes:
                break
        return None

    def __next__(self):
        while True:
            col = self._next()
            if col is not None:
                return PileupColumn(col[0], col[1], col[2], self.min_bq)
            if self.is_eof:
                raise StopIteration
            try:
                b = next(self.src)
            except StopIteration:
                b = None
            self._push(b)


class FastaFile:
    def __init__(self, path):
        self.filename = path
        self.references, self._len, self._seq = [], {}, {}
        idx = {}
        for line in open(path + ".fai"):
            p = line.rstrip("\n").split("\t")
            self.references.append(p[0])
            idx[p[0]] = (int(p[1]), int(p[2]), int(p[3]), int(p[4]))
            self._len[p[0]] = int(p[1])
        self._idx = idx
        self._f = open(path, "rb")
        self.lengths = [self._len[r] for r in self.references]

    def get_reference_length(self, name):
        return self._len[name]

    def _contig(self, name):
        if name not in self._seq:
            ln, off, lb, lw = self._idx[name]
            self._f.seek(off)
            nlines = (ln + lb - 1) // lb
            raw = self._f.read(ln + nlines * (lw - lb))
            self._seq[name] = raw.replace(b"\n", b"").replace(b"\r", b"")[:ln].decode()
        return self._seq[name]

    def fetch(self, reference=None, start=None, end=None, region=None):
        if reference not in self._idx:
            raise KeyError("sequence '%s' not present" % reference)
        if start is not None and start < 0:
            raise ValueError("start out of range (%i)" % start)
        s = self._contig(reference)
        start = 0 if start is None else start
        end = len(s) if end is None else end
        if end < start:
            raise ValueError("end before start")
        return s[start:end]

    def close(self):
        self._f.close()
