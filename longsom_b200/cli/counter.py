"""Drop-in BaseCellCounter: same CLI, same TSV (reference: workflow/scripts/SNVCalling/BaseCellCounter.py).

The reference fans 50 kb windows out to a multiprocessing pool and walks pysam pileup columns in
Python (:182-320, :392-406).  Here the BAM is decoded once on the host, every window of the run
goes to the GPU in one batch (sharded over LONGSOM_GPUS devices), and the per-site counts come
back as one table.  --nprocs is accepted and ignored (there is no process pool)."""
import argparse
import os
import sys
import timeit

from .. import bamio
from ..engine import CountParams
from ..pipeline import (bind_near_gpu, count_sites, devices_from_env, load_bam_for_counting, prewarm, prune_and_sort_windows, read_ends,
                        stream_count, write_counter_tsv)
from ..windows import make_windows


def initialize_parser():
    # flag names, types and defaults of BaseCellCounter.py:323-342
    p = argparse.ArgumentParser(description='Script to obtain a list of base and cell counts in scRNA bam file')
    p.add_argument('--bam', type=str, default=1, help='BAM file to be analysed', required=True)
    p.add_argument('--ref', type=str, default=1, help='Path to reference genome version. *fai must be available in the same directory as the reference genome file', required=True)
    p.add_argument('--chrom', type=str, help='Chromosome to be analysed. --chrom all to analyse all chromosomes', required=True)
    p.add_argument('--out_folder', default='.', help='Out folder', required=False)
    p.add_argument('--id', help='Prefix used to name output file. If provided, please conform with the following format: *.[cell_type] . Example: sample1.t_cell. If not provided, the basename of the BAM file will be used.', required=False)
    p.add_argument('--nprocs', default=1, help='Number of processes [Default: 1] (accepted for compatibility; the GPU path has no process pool)', required=False, type=int)
    p.add_argument('--bin', type=int, default=50000, help='Bin size for running the analysis [Default: 50000]', required=False)
    p.add_argument('--bed', type=str, default='', help='Regions to focus the analysis on. Three-column bed file listing the chromosome, start and end for those regions to be analysed.', required=False)
    p.add_argument('--bed_out', type=str, default='', help='Regions to ignore in the analysis. Three-column bed file listing the chromosome, start and end for those regions to be ignored.', required=False)
    p.add_argument('--min_ac', type=int, default=0, help='Minimum number of reads supporting the alternative allele required to consider a genomic site for mutation calling. Default: 0', required=False)
    p.add_argument('--min_af', type=float, default=0, help='Minimum alternative allele fraction required to consider a genomic site for mutation calling. Default = 0', required=False)
    p.add_argument('--min_dp', type=int, default=5, help='Minimum depth of coverage required to consider a genomic site for mutation calling. Default: 5', required=False)
    p.add_argument('--min_cc', type=int, default=5, help='Minimum number of cells required to consider a genomic site for mutation calling. Default: 5', required=False)
    p.add_argument('--min_bq', type=int, default=20, help='Minimum base quality to compute allele counts. Default: 20', required=False)
    p.add_argument('--min_mq', type=int, default=255, help='Minimum mapping quality required to consider a read for analysis. Default: 255', required=False)
    p.add_argument('--tmp_dir', type=str, default='.', help='Path to a directory to be used to store temporary files during processing', required=False)
    return p


def run(args):
    ID = args.id
    if ID is None:
        ID = os.path.basename(args.bam).replace(".bam", "")
    out_file = args.out_folder + '/' + str(ID) + ".tsv"
    print("Outfile: ", out_file, "\n")
    if args.tmp_dir != '.':
        # the Snakemake rule declares tmp_dir as an output directory (SNVCalling.smk:36): it must exist afterwards
        try:
            os.mkdir(args.tmp_dir)
            print("Directory ", args.tmp_dir, " created\n")
        except FileExistsError:
            print("Directory ", args.tmp_dir, " already exists\n")
    else:
        print("Not temp directory specified, using working directory as temp")
    if args.min_dp < 1:
        raise SystemExit("longsom_b200 BaseCellCounter: --min_dp must be >= 1")
    if len(devices_from_env()) == 1:
        bind_near_gpu(devices_from_env()[0])  # host threads and pinned staging buffers on the GPU's NUMA node
    prewarm(devices_from_env())  # CUDA contexts come up while the BAM is decoded
    fa = bamio.Fasta(args.ref)
    named = make_windows(fa.references, fa.lengths, args.chrom, args.bin, args.bed, args.bed_out)
    params = CountParams(min_bq=args.min_bq, min_mq=args.min_mq, min_dp=args.min_dp, min_cc=args.min_cc,
                         min_ac=args.min_ac, max_depth=200000)
    if os.environ.get("LONGSOM_STREAM", "1") != "0":
        # default: the BAM streams through pinned staging buffers in chunks (bounded host memory, decode / device /
        # writer overlapped); LONGSOM_STREAM=0 decodes the whole file first (round-1 path)
        stream_count(args.bam, named, fa, params, out_file, ID, devices_from_env())
        fa.close()
        return
    bd, batch, _cells = load_bam_for_counting(args.bam)
    ends = read_ends(batch)
    iv = prune_and_sort_windows(named, bd.contig_names, batch, ends)
    contig_seq = {t: fa.contig(bd.contig_names[t]) for t in sorted({w[0] for w in iv})}
    params = CountParams(min_bq=args.min_bq, min_mq=args.min_mq, min_dp=args.min_dp, min_cc=args.min_cc,
                         min_ac=args.min_ac, max_depth=200000)
    sites = count_sites(batch, iv, contig_seq, params, devices_from_env())
    write_counter_tsv(out_file, ID, sites, bd.contig_names)
    fa.close()


def main(argv=None):
    args = initialize_parser().parse_args(argv)
    start = timeit.default_timer()
    run(args)
    stop = timeit.default_timer()
    print("Computation time: " + str(round(stop - start)) + ' seconds')
    if os.environ.get("LS_STREAM_TIMING"):
        from ..pipeline import _peak_rss_gb
        print("[counter] peak RSS %.2f GB" % _peak_rss_gb(), file=sys.stderr)


if __name__ == '__main__':
    main(sys.argv[1:])
