"""pybedtools shim -- TEST INFRASTRUCTURE (oracle).  Only what MakeWindows uses
(/root/reference/workflow/scripts/SNVCalling/BaseCellCounter.py:81-113): BedTool from a list of
tuples or a BED path, merge(d=), intersect, filter (items expose .chrom), subtract,
window_maker(b, w=) and iteration over intervals indexable as [0], [1], [2] (strings, like
pybedtools.Interval)."""


class Interval:
    __slots__ = ("chrom", "start", "end")

    def __init__(self, chrom, start, end):
        self.chrom, self.start, self.end = str(chrom), int(start), int(end)

    def __getitem__(self, i):
        return (self.chrom, str(self.start), str(self.end))[i]

    def __repr__(self):
        return "%s\t%d\t%d" % (self.chrom, self.start, self.end)


def _read_bed(path):
    out = []
    for line in open(path):
        if not line.strip() or line.startswith(("#", "track", "browser")):
            continue
        p = line.rstrip("\n").split("\t")
        out.append(Interval(p[0], p[1], p[2]))
    return out


class BedTool:
    def __init__(self, src):
        if isinstance(src, BedTool):
            self.iv = list(src.iv)
        elif isinstance(src, str):
            self.iv = _read_bed(src)
        else:
            self.iv = [x if isinstance(x, Interval) else Interval(x[0], x[1], x[2]) for x in src]

    def __iter__(self):
        return iter(self.iv)

    def __len__(self):
        return len(self.iv)

    def merge(self, d=0):
        out = []
        for x in self.iv:  # bedtools merge expects sorted input
            if out and out[-1].chrom == x.chrom and x.start <= out[-1].end + d:
                out[-1].end = max(out[-1].end, x.end)
            else:
                out.append(Interval(x.chrom, x.start, x.end))
        return BedTool(out)

    def intersect(self, b):
        out = []
        for x in self.iv:
            for y in BedTool(b).iv:
                if x.chrom == y.chrom:
                    s, e = max(x.start, y.start), min(x.end, y.end)
                    if s < e:
                        out.append(Interval(x.chrom, s, e))
        return BedTool(out)

    def filter(self, fn):
        return BedTool([x for x in self.iv if fn(x)])

    def subtract(self, b):
        bb = BedTool(b).iv
        out = []
        for x in self.iv:
            pieces = [(x.start, x.end)]
            for y in bb:
                if y.chrom != x.chrom:
                    continue
                nxt = []
                for s, e in pieces:
                    if y.end <= s or y.start >= e:
                        nxt.append((s, e))
                    else:
                        if s < y.start:
                            nxt.append((s, y.start))
                        if y.end < e:
                            nxt.append((y.end, e))
                pieces = nxt
            out.extend(Interval(x.chrom, s, e) for s, e in pieces)
        return BedTool(out)

    def window_maker(self, b=None, w=None, **kw):
        out = []
        for x in BedTool(b).iv:
            s = x.start
            while s < x.end:
                e = min(s + w, x.end)
                out.append(Interval(x.chrom, s, e))
                s = e
        return BedTool(out)
