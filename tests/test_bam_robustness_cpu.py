"""Truncated and corrupt BAM files must end in an error string from both native readers (whole-file ls_bam_read,
streaming ls_bams_*), never in an out-of-bounds read: every length field of BGZF and BAM is checked against the bytes
that are there (csrc/host/ls_bamread.cpp, ls_bamstream.cpp)."""
import os
import struct
import zlib

import numpy as np
import pytest

from longsom_b200 import bamio, synth


def _bgzf(data, level=6):
    out = []
    for lo in range(0, max(len(data), 1), 0xff00):
        out.append(bamio._bgzf_block(data[lo:lo + 0xff00], level))
    out.append(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0]))
    return b"".join(out)


def _raw_of(path):
    """inflate a BGZF file with zlib (test side only)"""
    data, out, p = open(path, "rb").read(), [], 0
    while p < len(data):
        xlen = struct.unpack_from("<H", data, p + 10)[0]
        bsize = struct.unpack_from("<H", data, p + 16)[0] + 1
        out.append(zlib.decompress(data[p + 12 + xlen:p + bsize - 8], -15))
        p += bsize
    return b"".join(out)


@pytest.fixture(scope="module")
def small_bam(tmp_path_factory):
    d = synth.generate(**synth.config("C1", scale=0.02))
    b = d.batch
    path = str(tmp_path_factory.mktemp("bam") / "x.bam")
    names = [synth.barcode_of(c) + "-1" for c in range(d.n_cells + d.n_extra_cells)]
    bamio.write_bam(path, d.contig_names, d.contig_lens, b, lambda i: None if b.cell[i] < 0 else names[b.cell[i]])
    return path, b.n_reads


def _stream_all(path, chunk=1 << 16):
    class Slot:
        def __call__(self, n, nc, nb):
            return {"tid": np.zeros(n, np.int32), "pos": np.zeros(n, np.int32), "flag": np.zeros(n, np.uint16),
                    "mapq": np.zeros(n, np.uint8), "cell": np.zeros(n, np.int32), "l_qseq": np.zeros(n, np.int32),
                    "cigar_off": np.zeros(n + 1, np.uint32), "base_off": np.zeros(n + 1, np.uint64),
                    "cigar": np.zeros(nc, np.uint32), "seq4": np.zeros(nb // 2 + 1, np.uint8), "qual": np.zeros(nb, np.uint8)}
    bs = bamio.BamStream(path)
    n = 0
    try:
        while True:
            got = bs.next_chunk(chunk, Slot())
            if got is None:
                return n
            n += got[1]
    finally:
        bs.close()


def test_intact_file_reads_the_same_both_ways(small_bam):
    path, n_reads = small_bam
    assert bamio.read_bam(path).batch.n_reads == n_reads
    assert _stream_all(path) == n_reads
    assert _stream_all(path, chunk=1 << 22) == n_reads


def test_truncated_files_are_errors(small_bam, tmp_path):
    path, _ = small_bam
    data = open(path, "rb").read()
    rng = np.random.default_rng(1)
    cuts = [10, 17, 27, 100, len(data) // 3, len(data) // 2, len(data) - 29, len(data) - 40] + \
           [int(x) for x in rng.integers(30, len(data) - 30, 12)]
    for cut in cuts:
        p = str(tmp_path / ("cut%d.bam" % cut))
        with open(p, "wb") as f:
            f.write(data[:cut])
        # a cut exactly at a member boundary leaves a valid (shorter) file: fewer reads, no error.  Anything else
        # must raise.  Either way: no crash, no out-of-bounds read.
        for reader in (lambda q: bamio.read_bam(q).batch.n_reads, _stream_all):
            try:
                reader(p)
            except (IOError, OSError):
                pass


def test_corrupt_records_are_errors(small_bam, tmp_path):
    path, _ = small_bam
    raw = bytearray(_raw_of(path))
    # first record: right after the header
    l_text = struct.unpack_from("<I", raw, 4)[0]
    q = 8 + l_text
    n_ref = struct.unpack_from("<I", raw, q)[0]
    q += 4
    for _ in range(n_ref):
        q += 8 + struct.unpack_from("<I", raw, q)[0]
    rec = q + 4   # past block_size
    cases = {
        "l_seq far beyond block_size": (rec + 16, struct.pack("<I", 1 << 28)),
        "n_cigar beyond block_size": (rec + 12, struct.pack("<H", 0xffff)),
        "block_size below the fixed part": (q, struct.pack("<I", 8)),
        "block_size beyond the file": (q, struct.pack("<I", 0x7fffffff)),
        "l_text beyond the file": (4, struct.pack("<I", 0x7ffffff0)),
    }
    for name, (off, val) in cases.items():
        bad = bytearray(raw)
        bad[off:off + len(val)] = val
        p = str(tmp_path / "bad.bam")
        with open(p, "wb") as f:
            f.write(_bgzf(bytes(bad)))
        with pytest.raises((IOError, OSError)):
            bamio.read_bam(p)
        with pytest.raises((IOError, OSError)):
            _stream_all(p)


def test_corrupt_bgzf_members_are_errors(small_bam, tmp_path):
    path, _ = small_bam
    data = bytearray(open(path, "rb").read())
    cases = {
        "xlen beyond the member": (10, struct.pack("<H", 0xfff0)),
        "BSIZE smaller than the header": (16, struct.pack("<H", 5)),
        "not gzip": (0, b"\x00\x00"),
        "ISIZE above 64 KiB": (struct.unpack_from("<H", data, 16)[0] + 1 - 4, struct.pack("<I", 1 << 20)),
        "deflate payload damaged": (40, bytes(16)),
    }
    for name, (off, val) in cases.items():
        bad = bytearray(data)
        bad[off:off + len(val)] = val
        p = str(tmp_path / "badz.bam")
        with open(p, "wb") as f:
            f.write(bytes(bad))
        with pytest.raises((IOError, OSError)):
            bamio.read_bam(p)
        with pytest.raises((IOError, OSError)):
            _stream_all(p)
