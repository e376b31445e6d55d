"""Drop-in MergeBaseCellCounts (reference: workflow/scripts/SNVCalling/MergeBaseCellCounts.py).

Host glue between the per-cell-type BaseCellCounter tables and BaseCellCalling.step1: a positional
merge on (chrom, pos): N file cursors advanced in lock-step (:116-204), one row per table in memory,
chromosomes in lexicographic order, positions ascending, 'NA' where a cell type lacks the site."""
import argparse
import collections
import glob
import os
import sys
import time
import timeit

from ..pipeline import COUNTER_CONCEPTS

HEADER_LINES = 9  # ##fileDate + 7 ##INFO + #CHROM of a BaseCellCounter table (MergeBaseCellCounts.py:131)


def _most_common_join(values):
    """sort_set of the reference (:48-57): drop 'NA', drop '.' when something else is present,
    order by frequency (ties keep first-seen order), join with '|'."""
    vals = [v for v in values if v != 'NA']
    while len(vals) > 1 and '.' in vals:
        vals.remove('.')
    counter = collections.Counter(vals)
    return '|'.join(k for k, _ in sorted(counter.items(), key=lambda kv: kv[1], reverse=True))


class _Cursor:
    """One input table: the row its cursor points at.  A row whose position does not increase inside its chromosome
    is skipped, like updatefields of the reference (:8-23); past the last row chrom is '' (:177-180)."""
    __slots__ = ("f", "chrom", "pos", "ref", "info", "bc", "done")

    def __init__(self, path):
        self.f = open(path)
        for _ in range(HEADER_LINES):
            self.f.readline()
        self.chrom, self.pos, self.ref, self.info, self.bc, self.done = 'x', 0, 0, 0, 0, False
        first = self.f.readline()
        if first.strip() == "":
            self.done = True   # (the reference cannot parse a table without rows; here it contributes 'NA' everywhere)
            self.chrom, self.pos = '', -1
        else:
            self._take(first.strip())

    def _take(self, line):
        chrom, pos, ref, info, bc = line.split('\t')
        pos = int(pos)
        if chrom == self.chrom and self.pos >= pos:
            return
        self.chrom, self.pos, self.ref, self.info, self.bc = chrom, pos, ref, info, bc

    def advance(self):
        line = self.f.readline().strip()
        if line == "":   # end of file or a blank line: the table ends here
            self.done = True
            self.chrom, self.pos = '', -1
            return False
        self._take(line)
        return True


def merge_cell_types_files(infiles, outfile):
    """N file cursors advanced in lock-step over the (chrom, pos)-sorted tables, one row per file in memory
    (MergeBaseCellCounts.py:116-204): chromosomes in lexicographic order, positions ascending, 'NA' where a cell type
    lacks the site."""
    header = ['#CHROM', 'Start', 'End', 'REF', 'INFO']
    for path in infiles:
        header.append(os.path.basename(path).split('.')[-2])
    head_text = "##fileDate=%s\n" % time.strftime("%d/%m/%Y") + COUNTER_CONCEPTS + '\n' + '\t'.join(header) + '\n'
    if os.environ.get("LONGSOM_MERGE_NATIVE", "1") != "0":
        # the same cursor loop in native code (csrc/host/ls_merge.cpp); rows its strict parser refuses leave the table to
        # the Python loop below, whose exceptions are the reference's
        import ctypes as C
        from .. import bamio
        host = bamio._load_host()
        host.ls_merge_tables.restype = C.c_int
        host.ls_merge_tables.argtypes = [C.c_int32, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
        paths = (C.c_char_p * max(1, len(infiles)))(*[os.fsencode(p) for p in infiles])
        err = C.create_string_buffer(256)
        rc = host.ls_merge_tables(len(infiles), paths, os.fsencode(outfile), head_text.encode(), HEADER_LINES, err, 256)
        if rc == 0:
            return
        if rc < 0:
            raise IOError("MergeBaseCellCounts: " + err.value.decode())
    cursors = [_Cursor(path) for path in infiles]
    cur_chr, cur_pos = 1, 0
    with open(outfile, 'w') as out:
        out.write(head_text)
        while not all(c.done for c in cursors):
            for c in cursors:
                while c.chrom == cur_chr and c.pos <= cur_pos:
                    if not c.advance():
                        break
            go = not all(c.done for c in cursors)
            if any(c.chrom == cur_chr for c in cursors):
                if go:
                    cur_pos = min(c.pos for c in cursors if c.chrom == cur_chr and c.pos > cur_pos)
                    here = [c.chrom == cur_chr and c.pos == cur_pos for c in cursors]
                    refs = [str(c.ref) if h else 'NA' for c, h in zip(cursors, here)]
                    cells = [str(c.bc) if h else 'NA' for c, h in zip(cursors, here)]
                    # the INFO column joins the format strings of ALL cursors, on this site or not (:80)
                    fmt = _most_common_join([str(c.info) if c.info != 0 else 'NA' for c in cursors])
                    out.write('\t'.join([str(cur_chr), str(cur_pos), str(cur_pos), _most_common_join(refs), fmt]) + '\t' +
                              '\t'.join(cells) + '\n')
            elif go:
                cur_chr = sorted(c.chrom for c in cursors if c.chrom)[0]
                cur_pos = 0
    for c in cursors:
        c.f.close()


def initialize_parser():
    p = argparse.ArgumentParser(description='Script to merge the cell/base counts tsv files per cell type in only one')
    p.add_argument('--tsv_folder', type=str, default=1, help='Path to the directory containing the base count files in tsv format for each cell type. All tsv files in the directory will be used. Avoid not desired tsv files in this folder', required=True)
    p.add_argument('--outfile', help='Output file name', required=True)
    return p


def main(argv=None):
    args = initialize_parser().parse_args(argv)
    start = timeit.default_timer()
    print('-----------------------------------------------------------')
    print('1. Merging cell types in a unique tsv file')
    print('-----------------------------------------------------------\n')
    infiles = glob.glob(args.tsv_folder + '/*.tsv')  # same (filesystem) order as the reference => same column order
    if len(infiles) < 1:
        raise RuntimeError('No tsv files found')
    print(str(len(infiles)) + ' tsv files found\n')
    merge_cell_types_files(infiles, args.outfile)
    print('Done...\n')
    print('Time: ' + str(round(timeit.default_timer() - start, 2)) + ' seconds')


if __name__ == '__main__':
    main(sys.argv[1:])
