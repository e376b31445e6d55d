"""gnomAD_DB(path, gnomad_version=).get_info_from_df(df, "AF") over a plain TSV
(chrom, pos, ref, alt, AF); rows not in the table -> NaN (the reference replaces NaN by 0)."""
import os

import numpy as np
import pandas as pd


class gnomAD_DB:
    def __init__(self, path, gnomad_version="v4"):
        self.table = {}
        if path and os.path.isfile(path):
            for line in open(path):
                if line.startswith("#") or not line.strip():
                    continue
                c, p, r, a, af = line.rstrip("\n").split("\t")[:5]
                self.table[(c, int(p), r, a)] = float(af)

    def get_info_from_df(self, df, column):
        vals = [self.table.get((str(c), int(p), str(r), str(a)), np.nan)
                for c, p, r, a in zip(df["chrom"], df["pos"], df["ref"], df["alt"])]
        return pd.Series(vals, index=df.index, dtype=float)
