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


def _by_chrom(iv):
    d = {}
    for c, s, e in iv:
        d.setdefault(c, []).append((s, e))
    return d


def _intersect(a, b):
    """bedtools intersect: for every interval of a, in order, its overlap with every interval of b (in b's order).
    b is indexed by chromosome with a running maximum of the ends, so an exome-sized --bed costs a bisection per
    interval instead of a scan of all of b."""
    import bisect
    idx = {}
    for c, lst in _by_chrom(b).items():
        order = sorted(range(len(lst)), key=lambda k: lst[k][0])
        starts = [lst[k][0] for k in order]
        maxend, m = [], -1
        for k in order:
            m = max(m, lst[k][1])
            maxend.append(m)
        idx[c] = (lst, order, starts, maxend)
    out = []
    for c, s, e in a:
        if c not in idx:
            continue
        lst, order, starts, maxend = idx[c]
        hi = bisect.bisect_left(starts, e)          # candidates start before e ...
        lo = bisect.bisect_right(maxend, s, 0, hi)  # ... and nothing before lo reaches past s
        for k in sorted(order[lo:hi]):              # b's own order, like bedtools
            s2, e2 = lst[k]
            l2, h2 = max(s, s2), min(e, e2)
            if l2 < h2:
                out.append((c, l2, h2))
    return out


def _subtract(a, b):
    """bedtools subtract: every interval of a minus the union of b, pieces left to right."""
    import bisect
    union = {}
    for c, lst in _by_chrom(b).items():
        m = []
        for s, e in sorted(lst):
            if e <= s:
                continue
            if m and s <= m[-1][1]:
                m[-1][1] = max(m[-1][1], e)
            else:
                m.append([s, e])
        union[c] = ([x[0] for x in m], [x[1] for x in m])
    out = []
    for c, s, e in a:
        if c not in union:
            out.append((c, s, e))
            continue
        us, ue = union[c]
        k = bisect.bisect_right(ue, s)  # first removed block that ends after s
        cur = s
        while k < len(us) and us[k] < e:
            if us[k] > cur:
                out.append((c, cur, us[k]))
            cur = max(cur, ue[k])
            k += 1
        if cur < e:
            out.append((c, cur, e))
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
