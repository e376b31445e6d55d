"""Drop-in SplitBamCellTypes (reference: workflow/scripts/PreProcessing/SplitBamCellTypes.py).

The step in front of the hot path: one coordinate-sorted BAM in, one BAM (+ .bai)
per cell type out, plus '<id>.report.txt'.  The record routing, tag filters, end trimming, BGZF
writing and indexing are native (csrc/host/ls_bamsplit.cpp, multi-threaded inflate / deflate); this
module mirrors the reference's CLI, its barcode-table handling (meta_to_dict, :18-37) and its report.

Output BAMs hold the same records, byte for byte (qualities of trimmed ends zeroed), in the same
order as the reference's; the compressed bytes differ because members are deflated in parallel
(zlib level 6 like htslib's default; LONGSOM_BAM_LEVEL=1 trades ~30 % larger files for less CPU time:
1 M long reads / 394 MB in about 11 s at level 6 on 8 cores, deflate-bound)."""
import argparse
import ctypes as C
import os
import sys
import timeit

import pandas as pd

from .. import bamio

REASONS = ((1, 'nM'), (2, 'nM_not_found'), (4, 'NH'), (8, 'NH_not_found'), (16, 'MAPQ'))


def meta_to_dict(txt, tissue):
    """(:18-37) barcode (text before the first '-') -> cell type (blanks -> '_'), and the cell types in
    first-appearance order.  Later rows win for a repeated barcode, as dict() does."""
    metadata = pd.read_csv(txt, delimiter="\t")
    clean_index = metadata['Index'].str.replace('-.*$', '', regex=True)
    clean_type = metadata['Cell_type'].str.replace(' ', '_', regex=True)
    if tissue is not None:
        clean_type = str(tissue.replace(" ", "_")) + '__' + clean_type.astype(str)
    return dict(zip(clean_index, clean_type)), list(clean_type.unique())


def split_bam(bam, txt, outdir, donor, tissue, max_NM, max_NH, min_MAPQ, n_trim):
    start = timeit.default_timer()
    table, cell_types = meta_to_dict(txt, tissue)
    if len(table) < 1:
        print('Warning: No cell barcodes found in the --meta file')
        sys.exit()
    host = bamio._load_host()
    host.ls_bam_split.restype = C.c_int
    type_id = {t: i for i, t in enumerate(cell_types)}
    out_paths = ["{}/{}.{}.bam".format(outdir, donor, t) for t in cell_types]
    barcodes = [str(b).encode() for b in table]
    blob = b"".join(barcodes)
    off = (C.c_uint32 * (len(barcodes) + 1))()
    acc = 0
    for i, b in enumerate(barcodes):
        off[i] = acc
        acc += len(b)
    off[len(barcodes)] = acc
    types = (C.c_int32 * len(barcodes))(*[type_id[t] for t in table.values()])
    paths = (C.c_char_p * len(out_paths))(*[os.fsencode(p) for p in out_paths])
    counters = (C.c_int64 * 36)()
    first_seen = (C.c_int64 * 32)()
    err = C.create_string_buffer(512)
    rc = host.ls_bam_split(os.fsencode(bam), C.c_int(len(out_paths)), paths, blob, off, types, C.c_int64(len(barcodes)),
                           C.c_int(int(min_MAPQ)), C.c_int(-1 if max_NM is None else int(max_NM)),
                           C.c_int(-1 if max_NH is None else int(max_NH)), C.c_int(int(n_trim)),
                           C.c_int(min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))), C.c_int(int(os.environ.get("LONGSOM_BAM_LEVEL", "6"))), counters,
                           first_seen, err, C.c_int(512))
    if rc != 0:
        raise RuntimeError("ls_bam_split(%s): %s" % (bam, err.value.decode(errors="replace")))

    # report (:176-184): fixed keys, then the filter reasons in the order they first occurred, then the run time
    report = {'Total_reads': counters[0], 'Pass_reads': counters[1], 'CB_not_found': counters[2],
              'CB_not_matched': counters[3]}
    seen = sorted((first_seen[m], m) for m in range(1, 32) if first_seen[m])
    for _, m in seen:
        report[';'.join(name for bit, name in REASONS if m & bit)] = counters[4 + m]
    report['Total_time'] = round(timeit.default_timer() - start, 2)
    pd.DataFrame([report]).to_csv("{}/{}.report.txt".format(outdir, donor), index=False, sep='\t')
    return report


def initialize_parser():
    p = argparse.ArgumentParser(description='Split alignment file into cell type specific BAMs')
    p.add_argument('--bam', type=str, default=1, help='BAM file to be analysed (Sorted by coordinate)', required=True)
    p.add_argument('--meta', type=str, default=1,
                   help='Metadata file mapping cell barcodes to cell type information', required=True)
    p.add_argument('--id', type=str, default='Sample', help='Sample ID', required=False)
    p.add_argument('--max_nM', type=int, default=None,
                   help='Maximum number of mismatches permitted to consider reads for analysis. By default, this filter '
                        'is switched off, although we recommed using --max_nM 5. If applied, this filter requires having '
                        'the nM tag in the bam file. [Default: Switched off]', required=False)
    p.add_argument('--max_NH', type=int, default=None,
                   help='Maximum number of alignment hits permitted to consider reads for analysis. By default, this '
                        'filter is switched off, although we recommend using --max_NH 1. This filter requires having the '
                        'NH tag in the bam file. [Default: Switched off]', required=False)
    p.add_argument('--min_MQ', type=int, default=255,
                   help='Minimum mapping quality required to consider reads for analysis. Set this value to 0 to switch '
                        'this filter off. --min_MQ 255 is recommended for RNA data, and --min_MQ 30 for DNA data. '
                        '[Default: 255]', required=False)
    p.add_argument('--n_trim', type=int, default=0,
                   help='Number of bases trimmed by setting the base quality to 0 at the beginning and end of each read '
                        '[Default: 0]', required=False)
    p.add_argument('--outdir', default='.', help='Out directory', required=False)
    return p


def main(argv=None):
    args = initialize_parser().parse_args(argv)
    split_bam(args.bam, args.meta, args.outdir, args.id, None, args.max_nM, args.max_NH, args.min_MQ, args.n_trim)


if __name__ == '__main__':
    main(sys.argv[1:])
