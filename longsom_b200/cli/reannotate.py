"""Drop-in CellTypeReannotation (reference: workflow/scripts/CellTypeReannotation/CellTypeReannotation.py).

Host-only consumer of the HCCV genotyping table (K1' output): a cell is re-labelled
'Cancer' when, among the high-confidence variants it covers (VAF != '.'), the mutated fraction
(PASS calls + distinct fusion hits) reaches min_frac; cells covering fewer than min_variants sites
are dropped from the barcode table (collect_cells_with_SNVs :6-21, collect_cancer_cells :37-50,
write_reannotated_cell_types :52-69).  The barcode table goes through pandas read_csv / to_csv as in
the reference so that any extra columns keep their formatting."""
import argparse
import sys
import timeit
from collections import Counter

import pandas as pd


def covered_and_mutated(snv_file, min_variants):
    """-> (CB of every PASS call in an eligible cell [with repeats], {CB: covered sites}, eligible CBs)."""
    calls = pd.read_csv(snv_file, sep='\t')
    # the reference also derives an INDEX column here (:8); it is never used, but building it raises on
    # non-string contig names, which is kept
    calls['#CHROM'] + ':' + calls['Start'].astype(str) + ':' + calls['ALT_expected'].str.split(',', n=1, expand=True)[0]
    covered = Counter(calls['CB'][calls['VAF'] != '.'])
    eligible = [cb for cb, n in covered.items() if n >= min_variants]
    keep = calls['CB'].isin(eligible) & (calls['MutationStatus'] == 'PASS')
    return list(calls['CB'][keep]), covered, eligible


def fusion_cells(fusion_file):
    """One entry per distinct (fusion, barcode) pair, in last-occurrence order (:23-34)."""
    fusions = pd.read_csv(fusion_file, sep='\t')
    pairs = fusions['#FusionName'] + ':' + fusions['BC']
    return list(fusions['BC'][~pairs.duplicated(keep='last')])


def cancer_cells(snv_hits, fusion_hits, covered, min_variants, min_frac):
    hits = Counter(snv_hits + fusion_hits)
    called = []
    for cb, n in hits.items():
        frac = n / covered[cb] if covered[cb] >= min_variants else 0
        if frac >= min_frac:
            called.append(cb)
    return called


def write_reannotated_cell_types(cancer, eligible, bc_file, out_file):
    bcs = pd.read_csv(bc_file, sep='\t')
    bcs = bcs[bcs['Index'].isin(eligible)]
    cancer = set(cancer)
    bcs['Before_Reannotation_cell_type'] = bcs['Cell_type']
    bcs['Reannotated_cell_type'] = ['Cancer' if i in cancer else 'Non-Cancer' for i in bcs['Index']]
    bcs['Cell_type'] = bcs['Reannotated_cell_type']     # SplitBamCellTypes reads this column
    bcs.to_csv(out_file, sep='\t', index=False)


def initialize_parser():
    p = argparse.ArgumentParser(description='Script to get the alleles observed in each unique cell for the variant sites')
    p.add_argument('--SNVs', type=str, help='HCCV SNV calls (obtained by SingleCellGenotype.py), tsv file', required=True)
    p.add_argument('--fusions', type=str, help='HCCV fusion calls (obtained by CTAT_Fusion.smk), tsv file', required=True)
    p.add_argument('--outfile', type=str, help='Output tsv file', required=True)
    p.add_argument('--meta', type=str, help='Barcodes tsv file', required=True)
    p.add_argument('--min_variants', type=int, default=3, help='Minimum # of variants covered to consider a cell',
                   required=False)
    p.add_argument('--min_frac', type=float, default=0.2,
                   help='Minimum fraction of covered variants being mutated to call a cancer cell', required=False)
    return p


def main(argv=None):
    start = timeit.default_timer()
    args = initialize_parser().parse_args(argv)
    print("Outfile: ", args.outfile, "\n")
    snv_hits, covered, eligible = covered_and_mutated(args.SNVs, args.min_variants)
    fusion_hits = fusion_cells(args.fusions) if args.fusions else []
    cancer = cancer_cells(snv_hits, fusion_hits, covered, args.min_variants, args.min_frac)
    write_reannotated_cell_types(cancer, eligible, args.meta, args.outfile)
    print("Computation time: " + str(round(timeit.default_timer() - start)) + ' seconds')


if __name__ == '__main__':
    main(sys.argv[1:])
