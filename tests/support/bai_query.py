"""Independent .bai reader + region query (SAM spec 5.2 / 5.3) used to check the index the native splitter
writes: the records a samtools-style query returns for a region must equal a brute-force scan."""
import struct
import zlib


def read_bai(path):
    buf = open(path, "rb").read()
    assert buf[:4] == b"BAI\1"
    n_ref = struct.unpack_from("<i", buf, 4)[0]
    p = 8
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", buf, p)[0]
        p += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", buf, p)
            p += 8
            bins[b] = [struct.unpack_from("<QQ", buf, p + 16 * i) for i in range(n_chunk)]
            p += 16 * n_chunk
        n_intv = struct.unpack_from("<i", buf, p)[0]
        p += 4
        lin = list(struct.unpack_from("<%dQ" % n_intv, buf, p))
        p += 8 * n_intv
        refs.append((bins, lin))
    n_no_coor = struct.unpack_from("<Q", buf, p)[0] if p + 8 <= len(buf) else None
    return refs, n_no_coor


def reg2bins(beg, end):
    end -= 1
    out = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        out.extend(range(base + (beg >> shift), base + (end >> shift) + 1))
    return out


class BgzfReader:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        self.cache = {}

    def block(self, coff):
        if coff not in self.cache:
            b = self.buf
            assert b[coff:coff + 4] == b"\x1f\x8b\x08\x04", "virtual offset does not point at a BGZF member"
            xlen = struct.unpack_from("<H", b, coff + 10)[0]
            bsize = None
            x = coff + 12
            while x < coff + 12 + xlen:
                si1, si2, slen = struct.unpack_from("<BBH", b, x)
                if (si1, si2) == (66, 67):
                    bsize = struct.unpack_from("<H", b, x + 4)[0] + 1
                x += 4 + slen
            data = zlib.decompress(b[coff + 12 + xlen:coff + bsize - 8], -15)
            self.cache[coff] = (data, bsize)
        return self.cache[coff]

    def read(self, voff, n):
        """n bytes starting at virtual offset voff -> (bytes, next virtual offset)."""
        coff, uoff = voff >> 16, voff & 0xffff
        out = b""
        while n:
            data, bsize = self.block(coff)
            take = data[uoff:uoff + n]
            out += take
            n -= len(take)
            uoff += len(take)
            if uoff >= len(data):
                coff, uoff = coff + bsize, 0
                if n and coff >= len(self.buf):
                    raise EOFError
        return out, (coff << 16) | uoff


def query(bam_path, tid, beg, end):
    """(name, pos) of the records overlapping [beg, end) on reference tid, found through the index only."""
    refs, _ = read_bai(bam_path + ".bai")
    bins, lin = refs[tid]
    min_off = lin[min(beg >> 14, len(lin) - 1)] if lin else 0
    chunks = sorted(c for b in reg2bins(beg, end) if b in bins and b != 37450 for c in bins[b] if c[1] > min_off)
    rd = BgzfReader(bam_path)
    hits = []
    for c0, c1 in chunks:
        v = max(c0, min_off) if c0 < min_off < c1 else c0
        while v < c1:
            head, v2 = rd.read(v, 4)
            bs = struct.unpack("<I", head)[0]
            rec, v = rd.read(v2, bs)
            rtid, pos, l_name, _mq, _bin, n_cig, flag, _lseq = struct.unpack_from("<iiBBHHHI", rec, 0)
            cig = struct.unpack_from("<%dI" % n_cig, rec, 32 + l_name)
            span = sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8))
            rend = pos + (span if span > 0 and not flag & 4 else 1)
            if rtid == tid and pos < end and rend > beg:
                hits.append((rec[32:32 + l_name - 1].decode(), pos))
            if rtid != tid or pos >= end:
                break
    return sorted(set(hits))
