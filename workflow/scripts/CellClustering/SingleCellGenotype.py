#!/usr/bin/env python
"""Drop-in replacement of the reference's workflow/scripts/CellClustering/SingleCellGenotype.py (same CLI, same outputs),
backed by the B200-native longsom_b200 package.  Copy this workflow/ tree over the reference's, or
point the Snakemake rules' script path here; nothing else in workflow/rules changes."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
if __name__ == '__main__':
    from longsom_b200 import _early  # noqa: E402
    _early.warm()  # the CUDA context comes up while the numeric stack below is imported
from longsom_b200.cli.genotype import main  # noqa: E402

if __name__ == '__main__':
    main(sys.argv[1:])
