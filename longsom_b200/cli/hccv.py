"""Drop-in HCCVSingleCellGenotype (reference: workflow/scripts/CellTypeReannotation/
HCCVSingleCellGenotype.py): the genotype kernel path with the three behavioural differences
of that script (raw-CB lookup :163-169, VAF only when ALT > 0 :193-212, 14 columns and an
exact --outfile path :214-216,296-297).  See cli/genotype.py."""
import sys

from .genotype import main as _main


def main(argv=None):
    _main(argv, hccv=True)


if __name__ == '__main__':
    main(sys.argv[1:])
