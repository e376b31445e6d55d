"""Genomic window grid of the reference's MakeWindows (BaseCellCounter.py:81-113), as plain host
interval arithmetic (the reference shells out to bedtools through pybedtools):

  contigs as (name, 1, len)  [reference position 0 is never analysed, :86]
  --bed      -> merge(d=1) then intersect with the contigs          (:90-94)
  --chrom    -> keep one contig unless 'all'                        (:97-100)
  --bed_out  -> subtract                                            (:103-107)
  makewindows -w bin                                                (:110)
"""


def _read_bed(path):
    out = []
    with open(path) as f:
        for line in f:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            p = line.rstrip("\n").split("\t")
            out.append((p[0], int(p[1]), int(p[2])))
    return out


def _merge(iv, d):
    out = []
    for c, s, e in iv:  # bedtools merge expects sorted input, like the reference
        if out and out[-1][0] == c and s <= out[-1][2] + d:
            out[-1][2] = max(out[-1][2], e)
        else:
            out.append([c, s, e])
    return [tuple(x) for x in out]


def _intersect(a, b):
    out = []
    for c, s, e in a:
        for c2, s2, e2 in b:
            if c == c2:
                lo, hi = max(s, s2), min(e, e2)
                if lo < hi:
                    out.append((c, lo, hi))
    return out


def _subtract(a, b):
    out = []
    for c, s, e in a:
        pieces = [(s, e)]
        for c2, s2, e2 in b:
            if c2 != c:
                continue
            nxt = []
            for ps, pe in pieces:
                if e2 <= ps or s2 >= pe:
                    nxt.append((ps, pe))
                else:
                    if ps < s2:
                        nxt.append((ps, s2))
                    if e2 < pe:
                        nxt.append((e2, pe))
            pieces = nxt
        out.extend((c, ps, pe) for ps, pe in pieces)
    return out


def make_windows(contig_names, contig_lens, chrom="all", bin_size=50000, bed="", bed_out=""):
    """List of (chrom_name, start, end), 0-based half-open, in the order the reference generates them."""
    a = [(n, 1, int(l)) for n, l in zip(contig_names, contig_lens)]
    if bed != "":
        a = _intersect(_merge(_read_bed(bed), 1), a)
    if chrom != "all":
        a = [x for x in a if x[0] == chrom]
    if bed_out != "":
        a = _subtract(a, _read_bed(bed_out))
    out = []
    for c, s, e in a:
        while s < e:
            t = min(s + bin_size, e)
            out.append((c, s, t))
            s = t
    return out
