"""Drop-in BaseCellCalling.step2 (reference: workflow/scripts/SNVCalling/BaseCellCalling.step2.py).

GPU part (K3): the RNA-editing / PoN_SR / PoN_LR lookups.  The reference builds
{chrom: {pos // 20000: set(pos)}} dictionaries in Python (build_dict, :197-221 -- minutes for
1e7 rows) and probes them per candidate (:141-160); here each list becomes one uint64 key table
(contig id << 32 | pos), radix-sorted on the GPU and probed with one binary search per candidate.
Host part kept byte-compatible: the awk-style candidate prefilter (:23-29), the sliding 3-row
'Clustered' logic with its first/last-row handling (:59-92), label order, the pandas
read_csv/to_csv round trip and the gnomAD label (:96-117, :223-235)."""
import argparse
import os
import sys
import timeit

import numpy as np
import pandas as pd

from ..engine import Engine
from ..pipeline import devices_from_env


def read_site_list(path):
    """Parse an editing / PoN list: tab separated, '#' comments, columns chrom, pos.  ANY failure
    (missing file, gzip bytes, malformed row) yields an empty list -- the reference's bare
    `except` (:199-220) silently disables the filter in that case (quirk Q9)."""
    out = []
    try:
        with open(path, 'r') as f:
            for line in f:
                if not line.startswith('#'):
                    elements = line.split('\t')
                    out.append((elements[0], int(elements[1])))
    except Exception:
        out = []
    return out


def site_list_keys(path, key_ids):
    """uint64 keys (contig id << 32 | pos) of an editing / PoN list, read by the native parser
    (csrc/host/ls_sitelist.cpp); key_ids(names) -> contig ids shared with the candidates.  Lines the native parser
    refuses go through read_site_list, so that the reference's failure rule (Q9) decides."""
    import ctypes as C
    from .. import bamio
    if os.environ.get("LONGSOM_STEP2_NATIVE", "1") != "0":
        host = bamio._load_host()
        host.ls_sitelist_read.restype = C.c_int
        host.ls_sitelist_read.argtypes = [C.c_char_p, C.c_void_p]
        host.ls_sitelist_n.restype = C.c_int64
        host.ls_sitelist_n.argtypes = [C.c_void_p]
        host.ls_sitelist_n_chroms.restype = C.c_int32
        host.ls_sitelist_n_chroms.argtypes = [C.c_void_p]
        host.ls_sitelist_chrom.restype = C.c_char_p
        host.ls_sitelist_chrom.argtypes = [C.c_void_p, C.c_int32]
        host.ls_sitelist_fill.argtypes = [C.c_void_p] * 3
        host.ls_sitelist_free.argtypes = [C.c_void_p]
        h = C.c_void_p()
        rc = host.ls_sitelist_read(os.fsencode(path) if path else b"", C.byref(h))
        if rc == 0:
            try:
                n = host.ls_sitelist_n(h)
                chrom, pos = np.zeros(n, np.int32), np.zeros(n, np.int64)
                host.ls_sitelist_fill(h, chrom.ctypes.data, pos.ctypes.data)
                names = [host.ls_sitelist_chrom(h, i).decode() for i in range(host.ls_sitelist_n_chroms(h))]
            finally:
                host.ls_sitelist_free(h)
            gid = np.array(key_ids(names), np.uint64) if names else np.zeros(0, np.uint64)
            ok = (pos >= 0) & (pos < (1 << 32))
            return (gid[chrom[ok]] << np.uint64(32)) | pos[ok].astype(np.uint64)
        if rc < 0:
            return np.zeros(0, np.uint64)   # unreadable: the reference's bare except leaves the filter empty
    sites = read_site_list(path)
    ids = {}
    for c, _p in sites:
        if c not in ids:
            ids[c] = None
    names = list(ids)
    gid = dict(zip(names, key_ids(names)))
    return np.array([(gid[c] << 32) | p for c, p in sites if 0 <= p < (1 << 32)], np.uint64)


class GnomadTable:
    """Stand-in provider for gnomad_db.database.gnomAD_DB (step2.py:100-108) when that package is
    not installed: a TSV of chrom, pos, ref, alt, AF.  Unknown variants -> NaN (then 0)."""

    def __init__(self, path):
        self.table = {}
        if path and os.path.isfile(path):
            with open(path) as f:
                for line in f:
                    if line.startswith('#') or not line.strip():
                        continue
                    c, p, r, a, af = line.rstrip('\n').split('\t')[:5]
                    self.table[(c, int(p), r, a)] = float(af)

    def get_info_from_df(self, df, column):
        vals = [self.table.get((str(c), int(p), str(r), str(a)), np.nan)
                for c, p, r, a in zip(df['chrom'], df['pos'], df['ref'], df['alt'])]
        return pd.Series(vals, index=df.index, dtype=float)


def open_gnomad(path):
    try:
        from gnomad_db.database import gnomAD_DB  # the reference's provider, when available
        return gnomAD_DB(path, gnomad_version="v4")
    except ImportError:
        return GnomadTable(path)


def gnomad_label(FILTER, VAF, max_vaf):
    if FILTER == 'PASS':
        return 'gnomAD' if VAF >= max_vaf else 'PASS'
    return FILTER + ',gnomAD' if VAF >= max_vaf else FILTER


def variant_calling_step2(infile, distance, editing, pon_SR, pon_LR, gnomAD_db, max_gnomAD_VAF, outfile, engine):
    comments, column_names, cands = [], None, []
    with open(infile) as f:
        for line in f:
            if line.startswith('#'):
                if '#CHROM' in line:
                    column_names = line.rstrip('\n').split('\t')
                else:
                    comments.append(line)
                continue
            # awk prefilter of the reference: keep rows with $5 != "." and $6 != ".".  Almost every row of a step1
            # table fails it, so only the first six columns are cut out before the row is split in full.
            head = line.rstrip('\n').split('\t', 6)
            if len(head) > 5 and head[4] != "." and head[5] != ".":
                cands.append(line.rstrip('\n').split('\t'))

    # ---- K3: membership of every candidate in the three site lists, on the GPU -----------------
    chrom_id = {}

    def key_of(chrom, pos):
        cid = chrom_id.setdefault(chrom, len(chrom_id))
        return (cid << 32) | (pos & 0xffffffff)
    qkeys = np.array([key_of(c[0], int(c[1])) for c in cands], np.uint64)
    hits = []
    for path in (editing, pon_SR, pon_LR):
        keys = site_list_keys(path, lambda names: [chrom_id.setdefault(n, len(chrom_id)) for n in names])
        hits.append(engine.site_mask(keys, qkeys) if len(cands) else np.zeros(0, np.uint8))
    EDIT, PSR, PLR = hits

    # ---- 'Clustered': neighbours within `distance` among the rows the reference would have in its
    # 3-row list when it scores the candidate (:59-92): both neighbours, except that the FIRST row of a
    # file with >= 3 rows only sees the two rows after it and the LAST row only the one before it --
    # which is the same set, because each row's list is [prev, row, next] clipped to the file.
    n = len(cands)
    pos = [int(c[1]) for c in cands]
    out_rows = []
    for i, cand in enumerate(cands):
        FILTER = cand[5]
        if FILTER != ".":
            if n >= 3:
                nb = [0, 1, 2] if i == 0 else ([n - 2, n - 1] if i == n - 1 else [i - 1, i, i + 1])
            else:
                nb = list(range(n))
            close = sum(1 for j in nb if cands[j][0] == cand[0] and pos[j] != pos[i] and abs(pos[j] - pos[i]) <= distance)
            e, sr, lr = bool(EDIT[i]), bool(PSR[i]), bool(PLR[i])
            if close > 0 or e or sr or lr:
                for flag, label in ((e, 'RNA_editing_db'), (close > 0, 'Clustered'), (sr, 'PoN_SR'), (lr, 'PoN_LR')):
                    if flag:
                        FILTER = label if FILTER == 'PASS' else FILTER + ',' + label
                cand = cand[:5] + [FILTER] + cand[6:]
        out_rows.append(cand)

    outfile_temp = outfile + '.temp'
    with open(outfile_temp, 'w') as f2:
        for r in out_rows:
            f2.write('\t'.join(r) + '\n')
    with open(outfile, 'w') as f3:
        f3.writelines(comments)
    # same pandas round trip as the reference (dtype inference and NA handling included)
    output_df = pd.read_csv(outfile_temp, sep='\t', comment='#', names=column_names)
    db = open_gnomad(gnomAD_db)
    gin = output_df[['#CHROM', 'Start', 'REF', 'ALT']].copy()
    gin.columns = ['chrom', 'pos', 'ref', 'alt']
    vaf = db.get_info_from_df(gin, "AF").replace(np.nan, 0)
    output_df['gnomAD_VAF'] = vaf
    output_df['FILTER'] = output_df.apply(lambda x: gnomad_label(x['FILTER'], x['gnomAD_VAF'], max_gnomAD_VAF), axis=1)
    output_df = output_df[column_names]
    output_df.to_csv(outfile, sep='\t', index=False, mode='a')
    os.remove(outfile_temp)
    return len(out_rows)


def initialize_parser():
    # flags of step2.py:237-248
    p = argparse.ArgumentParser(description='Script to perform the scRNA somatic variant calling')
    p.add_argument('--infile', type=str, help='Input file with all samples merged in a single tsv', required=True)
    p.add_argument('--outfile', type=str, help='Out file prefix', required=True)
    p.add_argument('--editing', type=str, help='RNA editing file to be used to remove RNA-diting sites', required=False)
    p.add_argument('--pon_SR', type=str, help='Short-read (SR) Panel of normals (PoN) file to be used to remove germline polymorphisms and recurrent artefacts', required=True)
    p.add_argument('--pon_LR', type=str, help='Long-read (LR) Panel of normals (PoN) file to be used to remove germline polymorphisms and recurrent artefacts', nargs='?', const='', required=False)
    p.add_argument('--min_distance', type=int, default=5, help='Minimum distance allowed between potential somatic variants [Default: 5]', required=False)
    p.add_argument('--gnomAD_db', type=str, help='gnomAD v4 database file', required=False)
    p.add_argument('--gnomAD_max', type=float, default=0.01, help='Maximum gnomAD population VAF [default 0.01]', required=False)
    return p


def main(argv=None):
    args = initialize_parser().parse_args(argv)
    start = timeit.default_timer()
    print('\n- Variant calling step 2\n')
    print("	> Editing file used: ", args.editing)
    print("	> PoN_SR file used: ", args.pon_SR)
    print("	> PoN_LR file used: ", args.pon_LR)
    outfile2 = args.outfile + '.calling.step2.tsv'
    with Engine(devices_from_env()[0]) as eng:
        variant_calling_step2(args.infile, args.min_distance, args.editing, args.pon_SR, args.pon_LR, args.gnomAD_db,
                              args.gnomAD_max, outfile2, eng)
    print('\nTotal computing time: ' + str(round(timeit.default_timer() - start, 2)) + ' seconds')


if __name__ == '__main__':
    main(sys.argv[1:])
