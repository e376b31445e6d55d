"""Drop-in HighConfidenceCancerVariants (reference:
workflow/scripts/CellTypeReannotation/HighConfidenceCancerVariants.py).

Host-only stage that selects, from a step2 table, the high-confidence cancer
variants used to re-annotate cells; its output feeds HCCVSingleCellGenotype (K1').  Like step3 it
only re-reads counts already produced by the GPU path, so there is no kernel here: what must match
is the filter order and the text of three files:

  <prefix>.HCCV.tsv   header comments + HCCV_FILTER line + the final table
  <prefix>.HCCV.tsv2  table after the depth filter  (appended, no comments: HCCV_SNV :48)
  <prefix>.HCCV.tsv3  table with the HCCV_FILTER verdict of every surviving row (:78)

Order (:30-86): INDEX from the ORIGINAL ALT -> drop 'Non-Cancer'-only rows -> multi-allelic collapse
(rows whose runner-up ALT has >= 5 % of the top ALT's cancer reads are deleted, :89-163) -> both
cell types must have DP >= min_dp, a missing column is 'NoCov' (:203-212) -> chrM rows keep only
the database-free ones, nuclear rows lose Noisy_site / LC_* / gnomAD / RNA_editing_db / PoN ->
VAF/MCF verdict (:215-259) -> PASS rows closer than clust_dist in (chrom, position-as-text) order
are dropped (:165-200, label 'Clust_dist<d>', chrM exempt)."""
import argparse
import sys
import timeit

import pandas as pd

from .step3 import BASES, _assign_rows, _cancer_slot, _contains_any, _info_counts, _tag

HCCV_LINE = ('##INFO=HCCV_FILTER,Description=Filter status of the variant site for cell reannotation '
             '(high-confidence cancer variants)\n')
REPLACED = ['ALT', 'FILTER', 'Cell_types', 'Bc', 'Cc', 'VAF', 'MCF', 'MultiAllelic_filter']


def collapse_or_delete(ref, alt, flt, ctypes_s, dp, nc, bc, cc, vaf, mcf, cancer_info, noncancer_info):
    """(:89-163) -> (ALT, FILTER, Cell_types, Bc, Cc, VAF, MCF, 'KEEP' | 'DELETE')."""
    if ref not in BASES:
        raise KeyError(ref)
    untouched = (alt, flt, ctypes_s, bc, cc, vaf, mcf)
    if 'Multi-allelic' not in flt and '|' not in alt:
        return untouched + ('KEEP',)
    ctypes = ctypes_s.split(',')
    two = len(ctypes) > 1
    if two:
        ic = _cancer_slot(ctypes)
    elif ctypes[0] != 'Cancer':
        return untouched + ('DELETE',)
    alt_reads = [int(v) for v in _info_counts(cancer_info, 3, 4)]
    alt_reads[BASES.index(ref)] = 0
    best = max(range(4), key=lambda k: (alt_reads[k], -k))     # first maximum, like numpy.argmax
    second = max(v for k, v in enumerate(alt_reads) if k != best)
    if not second / alt_reads[best] < 0.05:                     # ZeroDivisionError without ALT reads, as upstream
        return untouched + ('DELETE',)
    letter = BASES[best]
    bc_c = int(_info_counts(cancer_info, 3)[best])
    cc_c = int(_info_counts(cancer_info, 2)[best])
    clean = flt.replace('Multi-allelic,', '').replace(',Multi-allelic', '').replace('Multi-allelic', '')
    if not two:
        return letter, clean, ctypes_s, bc_c, cc_c, round(bc_c / int(dp), 4), round(cc_c / int(nc), 4), 'KEEP'
    inc = 1 - ic
    dps, ncs = dp.split(','), nc.split(',')
    bc_n = int(_info_counts(noncancer_info, 3)[best])
    cc_n = int(_info_counts(noncancer_info, 2)[best])
    vaf_c, mcf_c = round(bc_c / int(dps[ic]), 4), round(cc_c / int(ncs[ic]), 4)
    vaf_n, mcf_n = round(bc_n / int(dps[inc]), 4), round(cc_n / int(ncs[inc]), 4)
    return (letter + ',' + letter, clean, ctypes_s, '%d,%d' % (bc_n, bc_c), '%d,%d' % (cc_n, cc_c),
            '%s,%s' % (vaf_n, vaf_c), '%s,%s' % (mcf_n, mcf_c), 'KEEP')


def depth_verdict(info_a, info_b, min_dp):
    """(:203-212)"""
    if not isinstance(info_a, str) or not isinstance(info_b, str):
        return 'NoCov'
    if int(info_a.split('|')[0]) < min_dp or int(info_b.split('|')[0]) < min_dp:
        return 'LowDepth'
    return 'PASS'


def fraction_verdict(ctypes_s, vaf, mcf, min_dvaf, min_dmcf):
    """(:215-259)"""
    ctypes = ctypes_s.split(',')
    if len(ctypes) == 1:
        if ctypes[0] != 'Cancer':
            return 'NonCancer'
        return 'PASS' if float(vaf) >= min_dvaf and float(mcf) >= min_dmcf else 'Low VAF/MCF'
    ic = _cancer_slot(ctypes)
    vafs, mcfs = vaf.split(','), mcf.split(',')
    vaf_c, vaf_n = float(vafs[ic]), float(vafs[1 - ic])
    if vaf_c < 0.05:
        return 'NonSig'
    if vaf_n > 0.1 and vaf_c - vaf_n < 2 * min_dvaf:
        return 'Heterozygous'
    if vaf_n > 0.2:
        return 'Heterozygous'
    return 'LowDeltaMCF' if float(mcfs[ic]) - float(mcfs[1 - ic]) < min_dmcf else 'PASS'


def clustered_indices(index_col, clust_dist):
    """Same neighbour rule as step3's, but over every remaining row (:165-184)."""
    keys = sorted((tuple(ix.split(':')) for ix in index_col), key=lambda t: (t[0], t[1]))
    hit = set()
    for a, b in zip(keys, keys[1:]):
        if a[0] == b[0] and a[0] != 'chrM' and abs(int(a[1]) - int(b[1])) < clust_dist:
            hit.add(':'.join(a))
            hit.add(':'.join(b))
    return hit


def HCCV_SNV(SNVs, outfile, min_dp, deltaVAF, deltaMCF, clust_dist):
    comments, columns = [], None
    with open(SNVs) as f:
        for line in f:
            if not line.startswith('#'):
                break
            if '#CHROM' in line:
                columns = line.rstrip('\n').split('\t')
            else:
                comments.append(line)
    with open(outfile, 'w') as o:
        o.writelines(comments)
        o.write(HCCV_LINE)

    df = pd.read_csv(SNVs, sep='\t', comment='#', names=columns)
    df['INDEX'] = ['%s:%s:%s' % (c, s, a.split(',', 1)[0]) for c, s, a in zip(df['#CHROM'], df['Start'], df['ALT'])]
    df = df[df['Cell_types'] != 'Non-Cancer']
    collapsed = [collapse_or_delete(*row) for row in zip(
        df['REF'], df['ALT'], df['FILTER'], df['Cell_types'], df['Dp'], df['Nc'], df['Bc'], df['Cc'], df['VAF'],
        df['MCF'], df['Cancer'], df['Non-Cancer'])]
    _assign_rows(df, REPLACED, collapsed)
    df = df[df['MultiAllelic_filter'] == 'KEEP']
    df = df[columns + ['INDEX']]

    _assign_rows(df, 'DP_FILTER', [depth_verdict(a, b, min_dp) for a, b in zip(df['Cancer'], df['Non-Cancer'])])
    df = df[df['DP_FILTER'] == 'PASS']
    df.to_csv(outfile + '2', sep='\t', index=False, mode='a')

    is_mt = df['#CHROM'] == 'chrM'
    mt = df[is_mt].copy()
    df = df[~is_mt]
    mt = mt[~_contains_any(mt['FILTER'], ('Min', 'LR', 'gnomAD', 'LC', 'RNA'))] if len(mt) else mt
    df = df[~_contains_any(df['FILTER'], ('Noisy_site', 'LC_Upstream', 'LC_Downstream', 'gnomAD', 'RNA_editing_db',
                                          'PoN'))]
    df = pd.concat([df, mt])

    _assign_rows(df, 'HCCV_FILTER', [fraction_verdict(ct, v, m, deltaVAF, deltaMCF)
                                     for ct, v, m in zip(df['Cell_types'], df['VAF'], df['MCF'])])
    df.to_csv(outfile + '3', sep='\t', index=False, mode='a')
    df = df[df['HCCV_FILTER'] == 'PASS']

    label = 'Clust_dist%s' % clust_dist
    hit = clustered_indices(df['INDEX'], clust_dist)
    _assign_rows(df, 'FILTER', [(_tag(f, label) if ix in hit else f) for ix, f in zip(df['INDEX'], df['FILTER'])])
    df = df[~df['FILTER'].str.contains('dist', regex=True)]
    df.to_csv(outfile, sep='\t', index=False, mode='a')
    return df


def initialize_parser():
    p = argparse.ArgumentParser(description='Script to perform High-Confidence Cancer Variants calling')
    p.add_argument('--SNVs', type=str, help='', required=True)
    p.add_argument('--outfile', type=str, help='Out file prefix', required=True)
    p.add_argument('--min_dp', type=float, default=20, help='Minimum depth in both celltypes to call a HCCV',
                   required=True)
    p.add_argument('--deltaVAF', type=float, default=0.1, help='Delta VAF between cancer and non-cancer cells',
                   required=True)
    p.add_argument('--deltaMCF', type=float, default=0.4,
                   help='Delta MCF (cancer cell fraction) between cancer and non-cancer cells', required=True)
    p.add_argument('--clust_dist', type=int, default=10000,
                   help='Minimum distance required between two consecutive SNVs', required=False)
    return p


def main(argv=None):
    start = timeit.default_timer()
    args = initialize_parser().parse_args(argv)
    print('\n- High Confidence Cancer Variants calling\n')
    HCCV_SNV(args.SNVs, args.outfile + '.HCCV.tsv', args.min_dp, args.deltaVAF, args.deltaMCF, args.clust_dist)
    print('\nTotal computing time: ' + str(round(timeit.default_timer() - start, 2)) + ' seconds')


if __name__ == '__main__':
    main(sys.argv[1:])
