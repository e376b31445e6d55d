"""Drop-in BaseCellCalling.step3 (reference: workflow/scripts/SNVCalling/BaseCellCalling.step3.py).

Host-only stage that follows the GPU path: it turns step2's per-site table into
the final cancer-specific call set.  No pileup or statistics are recomputed here, so there is no
kernel; what has to match the reference is the filter order and the text of the two TSVs:

  1. rows called only in 'Non-Cancer' are dropped (:37);
  2. multi-allelic sites are collapsed onto the strongest cancer ALT, tagged 'Multi-Allelic' unless
     the runner-up has < 5 % of its reads (MultiAllelic_filtering, :163-231);
  3. chrM rows take their own route (:49-57, chrM_filtering :101-161): database/homopolymer/PoN
     rows are dropped and the rest are tested on depth >= 100 and deltaVAF / deltaMCF (two cell
     types) or VAF/MCF >= 0.05 (one cell type);
  4. nuclear rows: Min_cell_types rows dropped, LowDepth / NoCov on cancer BC/CC (:233-251),
     CancerNonSig / NonCancerSig from step1's per-cell-type verdicts (:254-280), then rows whose
     step2 FILTER mentions Noisy_site, LC_*, RNA_editing_db, PoN, Cell_type_noise or gnomAD dropped;
  5. PASS rows closer than clust_dist to the neighbour in the reference's (chrom, pos-as-text)
     ordering are tagged Clust_dist_<d> (tag_clustered_SNVs, :283-316; chrM exempt);
  6. '<prefix>.calling.step3.unfiltered.tsv' = everything left, '<prefix>.calling.step3.tsv' = PASS.

pandas is used for exactly what the reference uses it for -- read_csv's type inference and to_csv's
number formatting -- so that the files are byte-identical; the row logic is plain Python over
column lists."""
import argparse
import sys
import timeit

import pandas as pd

BASES = 'ACTG'
FINAL_FILTER_LINE = ('##INFO=FINAL_FILTER,Description=Final ilter status, including chrM contaminants '
                     'and clustered sites\n')
REPLACED = ['ALT', 'FILTER', 'Cell_types', 'Bc', 'Cc', 'VAF', 'MCF', 'STEP3FILTER']


def _tag(current, label):
    """Append a label to a STEP3FILTER value ('PASS' is replaced, anything else is comma-extended)."""
    return label if current == 'PASS' else current + ',' + label


def _cancer_slot(cell_types):
    """Index of 'Cancer' in a two-entry Cell_types list (the reference only looks at entries 0 and 1
    and raises on anything else, :104-109)."""
    if cell_types[0] == 'Cancer':
        return 0
    if cell_types[1] == 'Cancer':
        return 1
    raise UnboundLocalError("cannot access local variable 'i_Cancer' where it is not associated with a value")


def _info_counts(info, field, n=None):
    parts = info.split('|')[field].split(':')
    return parts if n is None else parts[:n]


def collapse_multiallelic(ref, alt, flt, ctypes_s, dp, nc, bc, cc, vaf, mcf, cancer_info, noncancer_info):
    """(:163-231) -> (ALT, FILTER, Cell_types, Bc, Cc, VAF, MCF, STEP3FILTER)."""
    i_ref = BASES.index(ref) if ref in BASES else None
    if i_ref is None:
        raise KeyError(ref)
    if 'Multi-allelic' not in flt and '|' not in alt:
        return alt, flt, ctypes_s, bc, cc, vaf, mcf, 'PASS'
    ctypes = ctypes_s.split(',')
    alt_reads = [int(v) for v in _info_counts(cancer_info, 3, 4)]
    alt_reads[i_ref] = 0
    best = max(range(4), key=lambda k: (alt_reads[k], -k))     # first maximum, like numpy.argmax
    top = alt_reads[best]
    second = max(v for k, v in enumerate(alt_reads) if k != best)
    # ZeroDivisionError when no ALT read exists in cancer, as in the reference (MAX2/MAX)
    verdict = 'PASS' if second / top < 0.05 else 'Multi-Allelic'
    letter = BASES[best]
    bc_c = int(_info_counts(cancer_info, 3)[best])
    cc_c = int(_info_counts(cancer_info, 2)[best])
    if len(ctypes) > 1:
        ic = _cancer_slot(ctypes)
        inc = 1 - ic
        dps, ncs = dp.split(','), nc.split(',')
        vaf_c = round(bc_c / int(dps[ic]), 4)
        mcf_c = round(cc_c / int(ncs[ic]), 4)
        bc_n = int(_info_counts(noncancer_info, 3)[best])
        cc_n = int(_info_counts(noncancer_info, 2)[best])
        vaf_n = round(bc_n / int(dps[inc]), 4)
        mcf_n = round(cc_n / int(ncs[inc]), 4)
        # written non-cancer first whatever the order in Cell_types (:196-200)
        return (letter + ',' + letter, flt, ctypes_s, '%d,%d' % (bc_n, bc_c), '%d,%d' % (cc_n, cc_c),
                '%s,%s' % (vaf_n, vaf_c), '%s,%s' % (mcf_n, mcf_c), verdict)
    flt = flt.replace('Multi-allelic,', '').replace(',Multi-allelic', '').replace('Multi-allelic', '')
    return letter, flt, ctypes_s, bc_c, cc_c, round(bc_c / int(dp), 4), round(cc_c / int(nc), 4), verdict


def chrm_verdict(current, ctypes_s, dp, vaf, mcf, min_dvaf, min_dmcf):
    """(:101-161)"""
    ctypes = ctypes_s.split(',')
    if len(ctypes) > 1:
        ic = _cancer_slot(ctypes)
        inc = 1 - ic
        d1, d2 = dp.split(',')
        if int(d1) < 100 or int(d2) < 100:
            return _tag(current, 'LowDepth')
        vafs, mcfs = vaf.split(','), mcf.split(',')
        if float(vafs[ic]) - float(vafs[inc]) < min_dvaf:
            return _tag(current, 'LowDeltaVAF')
        if float(mcfs[ic]) - float(mcfs[inc]) < min_dmcf:
            return _tag(current, 'LowDeltaMCF')
        return current
    if int(dp) < 100:
        return _tag(current, 'LowDepth')
    if float(vaf) < 0.05:
        return _tag(current, 'LowVAF')
    if float(mcf) < 0.05:
        return _tag(current, 'LowMCF')
    return current


def cancer_support_verdict(current, alt, cancer_info, min_reads, min_cells):
    """(:233-251) a missing Cancer column (NaN after read_csv) is 'NoCov'."""
    k = BASES.index(alt[0]) if alt[0] in BASES else None
    if k is None:
        raise KeyError(alt[0])
    if not isinstance(cancer_info, str):
        return _tag(current, 'NoCov')
    if int(_info_counts(cancer_info, 3)[k]) < min_reads or int(_info_counts(cancer_info, 2)[k]) < min_cells:
        return _tag(current, 'LowDepth')
    return current


def betabin_verdict(current, ctypes_s, ct_filter):
    """(:254-280)"""
    ctypes = ctypes_s.split(',')
    weak = ('Non-Significant', 'Low-Significance')
    if len(ctypes) == 1:
        return _tag(current, 'CancerNonSig') if ct_filter in weak else current
    ic = _cancer_slot(ctypes)
    verdicts = ct_filter.split(',')
    if verdicts[ic] in weak:
        return _tag(current, 'CancerNonSig')
    if verdicts[1 - ic] in ('PASS', 'Low-Significance'):
        return _tag(current, 'NonCancerSig')
    return current


def clustered_indices(index_col, step3_col, clust_dist):
    """INDEX values of PASS rows that sit closer than clust_dist to their neighbour in the reference's
    ordering: sorted on (chrom, position AS TEXT), chrM skipped (:283-301)."""
    keys = sorted((tuple(ix.split(':')) for ix, f in zip(index_col, step3_col) if f == 'PASS'),
                  key=lambda t: (t[0], t[1]))
    hit = set()
    for a, b in zip(keys, keys[1:]):
        if a[0] == b[0] and a[0] != 'chrM' and abs(int(a[1]) - int(b[1])) < clust_dist:
            hit.add(':'.join(a))
            hit.add(':'.join(b))
    return hit


def _assign_rows(df, cols, values):
    """df[cols] = per-row values.  On an EMPTY frame the reference's `df.apply(..., axis=1)` yields an empty
    DataFrame and the assignment (or the next .str call) fails inside pandas; the same pandas call is made here so
    that an input without any usable row ends the same way (non-zero exit, header-only outputs)."""
    if len(df) == 0:
        # pandas probes the row function with an empty Series; the reference's row functions raise on it
        # (e.g. alt_dict[ALT[0]] on NaN), and pandas then hands back a copy of the (empty) frame -- same outcome here
        probe = lambda x: x['__row_function_fails_on_the_probe__']
        if isinstance(cols, list):
            df[cols] = df.apply(probe, axis=1, result_type='expand')
        else:
            df[cols] = df.apply(probe, axis=1)
    elif isinstance(cols, list):
        df[cols] = pd.DataFrame(values, index=df.index, columns=cols)
    else:
        df[cols] = values


def _contains_any(series, words):
    return series.map(lambda s: any(w in s for w in words)).astype(bool)


def variant_calling_step3(infile, out_prefix, deltaVAF, deltaMCF, chrM_conta, min_ac_reads, min_ac_cells, clust_dist):
    final_path = out_prefix + '.calling.step3.tsv'
    unfiltered_path = out_prefix + '.calling.step3.unfiltered.tsv'
    comments, columns = [], None
    with open(infile) as f:
        for line in f:
            if not line.startswith('#'):
                break
            if '#CHROM' in line:
                columns = line.rstrip('\n').split('\t')
            else:
                comments.append(line)
    for path in (final_path, unfiltered_path):
        with open(path, 'w') as o:
            o.writelines(comments)
            o.write(FINAL_FILTER_LINE)

    df = pd.read_csv(infile, sep='\t', comment='#', names=columns)
    df = df[df['Cell_types'] != 'Non-Cancer']

    collapsed = [collapse_multiallelic(*row) for row in zip(
        df['REF'], df['ALT'], df['FILTER'], df['Cell_types'], df['Dp'], df['Nc'], df['Bc'], df['Cc'], df['VAF'],
        df['MCF'], df['Cancer'], df['Non-Cancer'])]
    _assign_rows(df, REPLACED, collapsed)
    df['INDEX'] = [('%s:%s:%s' % (c, s, a.split(',', 1)[0])) for c, s, a in zip(df['#CHROM'], df['Start'], df['ALT'])]

    is_mt = df['#CHROM'] == 'chrM'
    mt = df[is_mt].copy()
    df = df[~is_mt]
    mt = mt[~_contains_any(mt['FILTER'], ('Min', 'LR', 'gnomAD', 'LC', 'RNA'))] if len(mt) else mt
    if len(mt) > 0:
        mt['STEP3FILTER'] = [chrm_verdict(f, ct, dp, v, m, deltaVAF, deltaMCF) for f, ct, dp, v, m in
                             zip(mt['STEP3FILTER'], mt['Cell_types'], mt['Dp'], mt['VAF'], mt['MCF'])]

    df = df[~_contains_any(df['FILTER'], ('Min_cell_types',))]
    _assign_rows(df, 'STEP3FILTER', [cancer_support_verdict(f, a, ci, min_ac_reads, min_ac_cells)
                                     for f, a, ci in zip(df['STEP3FILTER'], df['ALT'], df['Cancer'])])
    _assign_rows(df, 'STEP3FILTER', [betabin_verdict(f, ct, cf) for f, ct, cf in
                                     zip(df['STEP3FILTER'], df['Cell_types'], df['Cell_type_Filter'])])
    df = df[~_contains_any(df['FILTER'], ('Noisy_site', 'LC_Upstream', 'LC_Downstream', 'RNA_editing_db', 'PoN',
                                          'Cell_type_noise', 'gnomAD'))]

    df = pd.concat([df, mt])
    label = 'Clust_dist_%s' % clust_dist
    hit = clustered_indices(df['INDEX'], df['STEP3FILTER'], clust_dist)
    _assign_rows(df, 'STEP3FILTER', [(_tag(f, label) if ix in hit else f)
                                     for ix, f in zip(df['INDEX'], df['STEP3FILTER'])])
    df.to_csv(unfiltered_path, sep='\t', index=False, mode='a')
    df = df[~df['STEP3FILTER'].str.contains('dist', regex=True)]
    df[df['STEP3FILTER'] == 'PASS'].to_csv(final_path, sep='\t', index=False, mode='a')
    return df


def initialize_parser():
    p = argparse.ArgumentParser(description='Script to perform the scRNA somatic variant calling')
    p.add_argument('--infile', type=str, help='Input file with all samples merged in a single tsv', required=True)
    p.add_argument('--outfile', type=str, help='Out file prefix', required=True)
    p.add_argument('--deltaVAF', type=float, default=0.3, help='Delta VAF between cancer and non-cancer cells',
                   required=True)
    p.add_argument('--deltaMCF', type=float, default=0.3,
                   help='Delta MCF (cancer cell fraction) between cancer and non-cancer cells', required=True)
    p.add_argument('--chrM_contaminant', type=str, default='True',
                   help='Use this option if chrM contaminants are observed in non-cancer cells', required=False)
    p.add_argument('--min_ac_reads', type=int, default=2, help='Minimum ALT reads', required=False)
    p.add_argument('--min_ac_cells', type=int, default=3, help='Minimum mutated cells', required=False)
    p.add_argument('--clust_dist', type=int, default=5,
                   help='Minimum distance required between two consecutive SNVs', required=False)
    return p


def main(argv=None):
    start = timeit.default_timer()
    args = initialize_parser().parse_args(argv)
    print('\n- Variant calling step 3\n')
    variant_calling_step3(args.infile, args.outfile, args.deltaVAF, args.deltaMCF, args.chrM_contaminant,
                          args.min_ac_reads, args.min_ac_cells, args.clust_dist)
    print('\nTotal computing time: ' + str(round(timeit.default_timer() - start, 2)) + ' seconds')


if __name__ == '__main__':
    main(sys.argv[1:])
