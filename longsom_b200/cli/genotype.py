"""Drop-in SingleCellGenotype / HCCVSingleCellGenotype (reference:
workflow/scripts/CellClustering/SingleCellGenotype.py and
workflow/scripts/CellTypeReannotation/HCCVSingleCellGenotype.py).

GPU part: K1' (ls_genotype_count) produces the dense Dp / Alt tensors [site, cell] for every
candidate site of the run in one pass over the reads, and K2 (ls_betabinom_sf) all
beta-binomial tails of the (site, cell) pairs with ALT > 0.  The reference does this with one
pysam pileup per 50 kb bin, a Python dict per (site, barcode) and one scipy call per pair
(:114-218).  Host part kept byte-compatible: barcode cleaning, bin codes and temp-file order,
Python set iteration order of the sites inside a bin (:112,130), row formatting, and the pandas
pivots with the natural-sorted index (:342-379)."""
import argparse
import math
import os
import re
import sys
import threading
import timeit

import numpy as np
import pandas as pd

from .. import bamio
from .._lib import CLASS_ID, LS_CLASS_NA
from ..engine import Engine
from ..pipeline import devices_from_env

_NUM = re.compile(r"(\d+)")


def natural_key(s):
    """natsort's default key: digit runs compare as integers, the rest as text."""
    parts = _NUM.split(str(s))
    return tuple(int(p) if i % 2 else p for i, p in enumerate(parts))


def meta_to_dict(meta_file, tissue):
    # SingleCellGenotype.py:230-250
    metadata = pd.read_csv(meta_file, delimiter="\t")
    metadata['Index_clean'] = metadata['Index'].str.replace('-.*$', '', regex=True)
    metadata['Cell_type_clean'] = metadata['Cell_type'].str.replace(' ', '_', regex=True)
    if tissue is not None:
        tissue = tissue.replace(" ", "_")
        metadata['Cell_type_clean'] = str(tissue) + '__' + metadata['Cell_type_clean'].astype(str)
    return metadata.set_index('Index_clean')['Cell_type_clean'].to_dict()


def build_dict_variants(variant_file, window):
    # SingleCellGenotype.py:253-274: bins keyed CHROM_floor(POS/window), insertion ordered
    d = {}
    with open(variant_file) as f:
        for line in f:
            if not line.startswith('#') and not line.startswith('Chr'):
                line = line.rstrip('\n')
                e = line.split('\t')
                code = e[0] + '_' + str(math.floor(int(e[1]) / float(window)))
                d.setdefault(code, []).append(line)
    return d


def genotype(args, hccv):
    if args.tmp_dir != '.':
        try:
            os.mkdir(args.tmp_dir)
            print("Directory ", args.tmp_dir, " created\n")
        except FileExistsError:
            print("Directory ", args.tmp_dir, " already exists\n")
    else:
        print("Not temp directory specified, using working directory as temp")
    meta_dict = meta_to_dict(args.meta, args.tissue)
    barcodes = list(meta_dict.keys())
    bc_index = {b: i for i, b in enumerate(barcodes)}
    bins = build_dict_variants(args.infile, args.bin)

    # ---- candidate sites of every bin -------------------------------------------------------------
    # the CUDA context comes up in the background while the BAM is decoded (the native decoder releases the GIL)
    warm = {}

    def _make_engine():
        try:
            warm["engine"] = Engine(devices_from_env()[0])
        except Exception as e:
            warm["error"] = e
    warm_thread = threading.Thread(target=_make_engine, daemon=True)
    warm_thread.start()
    bam = bamio.read_bam(args.bam)
    tid_of = {n: i for i, n in enumerate(bam.contig_names)}
    bin_info = []  # (chrom, Target_sites dict in file order)
    site_keys = {}
    for code, lines in bins.items():
        target = {}
        chrom = lines[0].split('\t')[0]
        for g in lines:
            g = g.split('\t')
            target[int(g[1]) - 1] = [g[3], g[4].split(',')[0], g[6], g[13]]
        bin_info.append((chrom, target))
        if chrom in tid_of:  # the reference would raise inside pysam for an unknown contig
            for pos0, v in target.items():
                site_keys[(tid_of[chrom], pos0)] = CLASS_ID.get(v[1], LS_CLASS_NA)
    keys = sorted(site_keys)
    site_tid = np.array([k[0] for k in keys], np.int32)
    site_pos = np.array([k[1] for k in keys], np.int32)
    alt_cls = np.array([site_keys[k] for k in keys], np.uint8)
    row_of = {k: i for i, k in enumerate(keys)}

    # ---- barcode -> column of the dense tensors ------------------------------------------------------
    if hccv:
        # HCCVSingleCellGenotype.py:163-169 looks the RAW tag up in the cleaned metadata keys
        raw_to_cell = np.array([bc_index.get(b, -1) for b in bam.barcodes], np.int32)
    else:
        raw_to_cell = np.array([bc_index.get(b.split("-")[0], -1) for b in bam.barcodes], np.int32)
    batch = bam.with_cells(raw_to_cell) if len(bam.barcodes) else bam.batch
    n_cells = len(barcodes)
    warm_thread.join()
    if "error" in warm:
        raise warm["error"]
    with warm["engine"] as eng:
        eng.upload(batch, None)
        # only the touched (site, cell) pairs come back, each with its beta-binomial tail already evaluated on the
        # device (ALT > 0, not the chrM shortcut); the dense rows of the reference are expanded from them below
        chrM_rows = np.array([bam.contig_names[t] == 'chrM' for t in site_tid], bool) if len(keys) else np.zeros(0, bool)
        skip = chrM_rows.astype(np.uint8) if args.chrM_contaminant == 'True' else None
        t_site, t_cell, t_dp, t_alt, t_p = eng.genotype_sparse(
            site_tid, site_pos, alt_cls, n_cells, args.alpha2, args.beta2, skip_p=skip, min_bq=args.min_bq,
            min_mq=args.min_mq, max_depth=200000, alt_only=(args.alt_flag != 'All'), bin_size=args.bin)
    t_p = np.round(t_p, 4)
    bounds = np.searchsorted(t_site, np.arange(len(keys) + 1))
    native = os.environ.get("LONGSOM_GENO_NATIVE", "1") != "0"
    touched = [None] * len(keys)
    if not native:
        # per site: {cell: (Dp, Alt, BetaBin)}
        tc, td, ta, tp = t_cell.tolist(), t_dp.tolist(), t_alt.tolist(), t_p.tolist()
        for r in range(len(keys)):
            lo_, hi_ = int(bounds[r]), int(bounds[r + 1])
            if hi_ > lo_:
                touched[r] = {tc[i]: (td[i], ta[i], tp[i]) for i in range(lo_, hi_)}
    else:
        import ctypes as C
        host = bamio._load_host()
        host.ls_geno_rows.restype = C.c_int64
        host.ls_geno_rows.argtypes = [C.c_int32] + [C.c_void_p] * 9 + [C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
        host.ls_geno_rows_free.argtypes = [C.c_void_p]
        t_cell = np.ascontiguousarray(t_cell, np.int32)
        t_dp = np.ascontiguousarray(t_dp, np.int32)
        t_alt = np.ascontiguousarray(t_alt, np.int32)
        t_p = np.ascontiguousarray(t_p, np.float64)
        cell_text = (C.c_char_p * max(1, n_cells))(*[(bc + '\t' + meta_dict[bc]).encode() for bc in barcodes])

    # ---- rows, in the reference's order ------------------------------------------------------------
    blocks = {}
    for chrom, target in bin_info:
        sites = set(target.keys())  # same construction as the reference => same iteration order
        lo, hi = min(sites), max(sites)
        if native:
            # the dense rows are expanded from the touched pairs by the native writer (csrc/host/ls_genorows.cpp)
            order = list(sites)  # CELLS dict comprehension iterates the set (:130)
            ns = len(order)
            prefix, index = [], []
            chrm = np.zeros(ns, np.uint8)
            hlo, hhi = np.zeros(ns, np.int64), np.zeros(ns, np.int64)
            for j, POS in enumerate(order):
                Ref_exp, Alt_exp, Cell_type_exp, Num_cells_exp = target[POS]
                prefix.append('\t'.join([str(chrom), str(POS + 1), str(POS + 1), Ref_exp, Alt_exp, str(Cell_type_exp),
                                         str(Num_cells_exp)]).encode())
                index.append((str(chrom) + ':' + str(POS + 1) + ':' + Alt_exp.split(',')[0]).encode())
                chrm[j] = 1 if (args.chrM_contaminant == 'True' and str(chrom) == 'chrM') else 0
                r = row_of.get((tid_of.get(chrom, -1), POS))
                if r is not None:
                    hlo[j], hhi[j] = bounds[r], bounds[r + 1]
            pre_a = (C.c_char_p * max(1, ns))(*prefix)
            idx_a = None if hccv else (C.c_char_p * max(1, ns))(*index)
            text = C.c_void_p()
            n = host.ls_geno_rows(ns, pre_a, idx_a, chrm.ctypes.data, hlo.ctypes.data, hhi.ctypes.data, t_cell.ctypes.data,
                                  t_dp.ctypes.data, t_alt.ctypes.data, t_p.ctypes.data, n_cells, cell_text, 1 if hccv else 0,
                                  float(args.pvalue), C.byref(text))
            if n < 0:
                raise RuntimeError("ls_geno_rows failed (%d)" % n)
            try:
                blocks[(str(chrom), lo)] = [C.string_at(text.value, n)] if n else []
            finally:
                if n >= 0 and text.value:
                    host.ls_geno_rows_free(text)
            continue
        rows = []
        for POS in sites:  # CELLS dict comprehension iterates the set (:130)
            Ref_exp, Alt_exp, Cell_type_exp, Num_cells_exp = target[POS]
            r = row_of.get((tid_of.get(chrom, -1), POS))
            hits = (touched[r] if r is not None else None) or {}
            for c, bc in enumerate(barcodes):
                CTYPE = meta_dict[bc]
                DP, ALT, PV = hits.get(c, (0, 0, None))
                VAF, BETABIN, MUTATED = '.', '.', 'NoCoverage'
                if DP > 0:
                    if not hccv:
                        VAF = round(ALT / DP, 4)
                    if ALT > 0:
                        if hccv:
                            VAF = round(ALT / DP, 4)
                        if args.chrM_contaminant == 'True' and str(chrom) == 'chrM':
                            MUTATED = 'LowVAFChrM' if VAF < 0.3 else 'PASS'
                        else:
                            BETABIN = np.float64(PV)
                            MUTATED = 'PASS' if BETABIN < args.pvalue else 'BetaBin_problem'
                    else:
                        if hccv:
                            VAF = float(0)
                        MUTATED = 'NoAltReads'
                group = [str(chrom), str(POS + 1), str(POS + 1), Ref_exp, Alt_exp, str(Cell_type_exp), str(Num_cells_exp),
                         bc, CTYPE, str(DP), str(ALT), str(VAF), str(BETABIN), str(MUTATED)]
                if not hccv:
                    BIN = 1 if MUTATED == "PASS" else (3 if MUTATED == "NoCoverage" else 0)
                    group += [str(BIN), str(chrom) + ':' + str(POS + 1) + ':' + Alt_exp.split(',')[0]]
                rows.append(('\t'.join(group) + '\n').encode())
        # temp file CHROM_min_max; a later bin with the same name overwrites an earlier one (:117-120)
        blocks[(str(chrom), lo)] = rows
    header = ['#CHROM', 'Start', 'End', 'REF', 'ALT_expected', 'Cell_type_expected', 'Num_cells_expected', 'CB',
              'Cell_type_observed', 'Dp', 'ALT', 'VAF', 'BetaBin', 'MutationStatus']
    if not hccv:
        header += ['BinMutationStatus', 'INDEX']
    long_path = args.outfile if hccv else args.outfile + '.SingleCellGenotype.tsv'
    if blocks:
        with open(long_path, 'wb') as out:
            out.write(('\t'.join(header) + '\n').encode())
            for chrom in sorted({k[0] for k in blocks}):
                for start in sorted(k[1] for k in blocks if k[0] == chrom):
                    out.writelines(blocks[(chrom, start)])
    else:
        print('No temporary files found')
    return long_path


def collect_cells_with_fusions(fusion_file):
    # SingleCellGenotype.py:325-340
    fusions = pd.read_csv(fusion_file, sep='\t')
    fusions['INDEX'] = fusions['#FusionName'] + ':' + fusions['BC']
    fusions = fusions.drop_duplicates(subset='INDEX', keep="last")
    lines = []
    for _, row in fusions.iterrows():
        lines.append(['.', '.', '.', '.', '.', '.', '.', row['BC'], '.', 1, 1, 1, '.', '.', 1, 'zzz:' + row['#FusionName']])
    return pd.DataFrame(lines)


def sort_chr_index(df):
    # chrM sorts last (renamed to chrZ for the natural sort), fusions ('zzz:') after everything (:342-348)
    df.index = [i.replace('chrM', 'chrZ') for i in df.index]
    df = df.reindex(sorted(df.index, key=natural_key))
    df.index = [i.replace('chrZ', 'chrM').replace('zzz:', '') for i in df.index]
    return df


def pivot_long_dataframe(out_prefix, fusions):
    long_df = pd.read_csv(out_prefix + '.SingleCellGenotype.tsv', sep='\t')
    if not fusions.empty:
        fusions.columns = long_df.columns
        long_df = pd.concat([long_df, fusions]).fillna(3)
    for values, name in (('Dp', 'DpMatrix'), ('ALT', 'AltMatrix'), ('VAF', 'VAFMatrix'), ('BinMutationStatus', 'BinaryMatrix')):
        m = long_df.pivot(index='INDEX', columns='CB', values=values)
        m = sort_chr_index(m)
        m.to_csv(out_prefix + '.' + name + '.tsv', sep='\t', index=True)


def initialize_parser(hccv):
    if hccv:
        p = argparse.ArgumentParser(description='Script to get the alleles observed in each unique cell for the variant sites')
    else:
        p = argparse.ArgumentParser(description='Script to get the SNV/fusions observed in each unique cell')
    p.add_argument('--bam', type=str, default=1, help='Tumor bam file to be analysed', required=True)
    p.add_argument('--infile', type=str, default=1, help='Base calling file (obtained by BaseCellCalling.step2.py), ideally only the PASS variants', required=True)
    p.add_argument('--ref', type=str, default=1, help='Reference genome. *fai must be available in the same folder as reference', required=True)
    p.add_argument('--meta', type=str, default=1, help='Metadata with cell barcodes per cell type', required=True)
    if not hccv:
        p.add_argument('--fusions', type=str, help='Fusions file from CTAT_fusion', nargs='?', const='', required=True)
    p.add_argument('--outfile', default='Matrix.tsv', help='Out file', required=False)
    p.add_argument('--alt_flag', default='All', choices=['Alt', 'All'], help='Flag to search for cells carrying the expected alt variant (Alt) or all cells independent of the alt allele observed (All)', required=False)
    p.add_argument('--nprocs', default=1, help='Number of processes [Default = 1] (accepted for compatibility)', required=False, type=int)
    p.add_argument('--bin', type=int, default=50000, help='Bin size for running the analysis [Default 50000]', required=False)
    p.add_argument('--min_bq', type=int, default=30, help='Minimum base quality permited for the base counts. Default = 30', required=False)
    p.add_argument('--min_mq', type=int, default=255, help='Minimum mapping quality required to analyse read. Default = 255', required=False)
    p.add_argument('--tissue', type=str, default=None, help='Tissue of the sample', required=False)
    p.add_argument('--tmp_dir', type=str, default='tmpDir', help='Temporary folder for tmp files', required=False)
    # the two scripts ship different defaults (SingleCellGenotype.py:396-397, HCCVSingleCellGenotype.py:332-333)
    p.add_argument('--alpha2', type=float, default=0.260288007167716 if hccv else 0.2474528917555431, help='Alpha parameter for Beta-binomial distribution of read counts.', required=False)
    p.add_argument('--beta2', type=float, default=173.94711910763732 if hccv else 162.03696139428595, help='Beta parameter for Beta-binomial distribution of read counts.', required=False)
    p.add_argument('--pvalue', type=float, default=0.01, help='P-value for the beta-binomial test to be significant', required=False)
    p.add_argument('--chrM_contaminant', type=str, default='True', help='Use this option if chrM contaminants are observed in non-cancer cells', required=False)
    return p


def main(argv=None, hccv=False):
    args = initialize_parser(hccv).parse_args(argv)
    start = timeit.default_timer()
    print("Outfile: " if hccv else "Outfile prefix: ", args.outfile, "\n")
    genotype(args, hccv)
    if not hccv:
        fusions = collect_cells_with_fusions(args.fusions) if args.fusions else pd.DataFrame()
        pivot_long_dataframe(args.outfile, fusions)
    print("Computation time: " + str(round(timeit.default_timer() - start)) + ' seconds')


if __name__ == '__main__':
    main(sys.argv[1:])
