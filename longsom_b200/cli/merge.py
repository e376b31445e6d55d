"""Drop-in MergeBaseCellCounts (reference: workflow/scripts/SNVCalling/MergeBaseCellCounts.py).

Host glue between the per-cell-type BaseCellCounter tables and BaseCellCalling.step1: a positional
merge on (chrom, pos).  The reference advances N file cursors in lock-step (:116-204); here every
table is indexed by (chrom, pos) and the union is emitted in the reference's order
(chrom lexicographic, pos ascending), with 'NA' where a cell type lacks the site."""
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


def merge_cell_types_files(infiles, outfile):
    tables, header = [], ['#CHROM', 'Start', 'End', 'REF', 'INFO']
    info_fmt = []  # last INFO format string seen per file (the reference passes this whole list to sort_set)
    for path in infiles:
        header.append(os.path.basename(path).split('.')[-2])
        rows = {}
        fmt = None
        with open(path) as f:
            for i, line in enumerate(f):
                if i < HEADER_LINES:
                    continue
                line = line.strip()
                if line == "":
                    break
                p = line.split('\t')
                rows[(p[0], int(p[1]))] = (p[2], p[3], p[4])
        tables.append(rows)
    keys = sorted(set().union(*[t.keys() for t in tables])) if tables else []
    # the reference keeps, per file, the fields of the line its cursor currently points at; the INFO
    # format column is the same constant on every line, so its joined set is that constant
    with open(outfile, 'w') as out:
        out.write("##fileDate=%s\n" % time.strftime("%d/%m/%Y"))
        out.write(COUNTER_CONCEPTS + '\n')
        out.write('\t'.join(header) + '\n')
        for t in tables:
            first = next(iter(t.values()), None)
            info_fmt.append(first[1] if first else 'NA')
        fmt_joined = _most_common_join(info_fmt)
        for chrom, pos in keys:
            refs, cells = [], []
            for t in tables:
                r = t.get((chrom, pos))
                if r is None:
                    refs.append('NA')
                    cells.append('NA')
                else:
                    refs.append(r[0])
                    cells.append(r[2])
            out.write('\t'.join([chrom, str(pos), str(pos), _most_common_join(refs), fmt_joined]) + '\t' +
                      '\t'.join(cells) + '\n')


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
