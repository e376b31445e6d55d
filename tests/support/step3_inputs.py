"""Fabricated step2-format table that drives every branch of BaseCellCalling.step3 (chrM routes with
deep coverage, both Cell_types orders, multi-allelic collapse in one and two cell types, missing
per-cell-type INFO, clustered PASS rows).  Used by oracle/make_golden.py to produce
tests/golden/s3/ from the unmodified reference script; the generated table itself is committed
there too, so the test does not depend on this generator staying bit-stable."""
import random

HEADER = (
    "##fileDate=01/01/2026\n"
    "##INFO=Cell_type_Filter,Description=Filter status of the variant site in each cell type\n"
)
COLUMNS = ["#CHROM", "Start", "End", "REF", "ALT", "FILTER", "Cell_types", "Up_context", "Down_context", "N_ALT", "Dp",
           "Nc", "Bc", "Cc", "VAF", "MCF", "BCp", "CCp", "Cell_types_min_BC", "Cell_types_min_CC", "Rest_BC", "Rest_CC",
           "Fisher_p", "Cell_type_Filter", "INFO", "Non-Cancer", "Cancer"]
BASES = "ACTG"
LABELS = ["Noisy_site", "LC_Upstream", "LC_Downstream", "RNA_editing_db", "PoN_SR", "PoN_LR", "Cell_type_noise", "gnomAD",
          "Min_cell_types", "Multiple_cell_types"]
VERDICTS = ["PASS", "PASS", "PASS", "Non-Significant", "Low-Significance"]


def _info(rng, depth, ref, alts):
    """DP|NC|CC|BC|BQ|BCf|BCr with six classes A:C:T:G:I:D; alts = [(base index, reads)]."""
    bc = [0] * 6
    for k, n in alts:
        bc[k] = n
    bc[BASES.index(ref)] = max(depth - sum(bc), 0)
    cc = [min(v, max(1, v * 2 // 3)) if v else 0 for v in bc]
    fwd = [v // 2 for v in bc]
    rev = [v - f for v, f in zip(bc, fwd)]
    bq = [v * 38 for v in bc]
    j = lambda xs: ":".join(str(x) for x in xs)
    return "%d|%d|%s|%s|%s|%s|%s" % (sum(bc), max(1, sum(cc)), j(cc), j(bc), j(bq), j(fwd), j(rev)), bc, cc


def rows(seed=20260101, n=420):
    rng = random.Random(seed)
    out = []
    pos = {"chr1": 1000, "chr2": 500, "chrM": 100}
    for i in range(n):
        chrom = rng.choice(["chr1", "chr1", "chr2", "chrM"])
        # short hops so that some PASS rows fall within clust_dist; 4/5-digit mix exercises the text ordering
        pos[chrom] += rng.choice([1, 3, 7, 40, 900, 2500])
        start = pos[chrom]
        ref = rng.choice(BASES)
        others = [b for b in BASES if b != ref]
        kind = rng.random()
        ctypes = ("Cancer" if kind < 0.55 else "Non-Cancer" if kind < 0.63 else
                  "Non-Cancer,Cancer" if kind < 0.85 else "Cancer,Non-Cancer")
        names = ctypes.split(",")
        multi = rng.random() < 0.22
        a1 = rng.choice(others)
        a2 = rng.choice([b for b in others if b != a1])
        deep = chrom == "chrM" or rng.random() < 0.3
        info, per = {}, {}
        for ct in ("Cancer", "Non-Cancer"):
            depth = rng.randrange(90, 3000) if deep else rng.randrange(4, 120)
            n1 = rng.randrange(1, max(2, depth // 2)) if (ct in names or rng.random() < 0.3) else 0
            if ct == "Non-Cancer" and rng.random() < 0.5:
                n1 = n1 // 20
            n2 = 0
            if multi:
                n2 = rng.choice([0, 1, max(1, n1 // 30), max(1, n1 // 3), n1])   # includes exact ties and < 5 % cases
            s, bc, cc = _info(rng, depth, ref, [(BASES.index(a1), n1), (BASES.index(a2), n2)])
            info[ct], per[ct] = s, (sum(bc), max(1, sum(cc)), bc[BASES.index(a1)], cc[BASES.index(a1)])
        missing = rng.random() < 0.12
        if len(names) == 1 and not multi and missing:
            other = "Non-Cancer" if names[0] == "Cancer" else "Cancer"
            info[other] = "NA"
        labels = [l for l in LABELS if rng.random() < 0.06]
        if multi:
            labels.insert(rng.randrange(len(labels) + 1), "Multi-allelic")
        flt = ",".join(labels) if labels else "PASS"
        alt_one = (a1 + "|" + a2) if multi and rng.random() < 0.7 else a1
        f = lambda xs: ",".join(str(x) for x in xs)
        dp = f(per[c][0] for c in names)
        nc = f(per[c][1] for c in names)
        bcs = f(per[c][2] for c in names)
        ccs = f(per[c][3] for c in names)
        vaf = f(round(per[c][2] / per[c][0], 4) for c in names)
        mcf = f(round(per[c][3] / per[c][1], 4) for c in names)
        ctf = f((rng.choice(VERDICTS) if not multi else rng.choice(VERDICTS + ["Multi-allelic"])) for _ in names)
        out.append([chrom, start, start, ref, f(alt_one for _ in names), flt, ctypes, "ACGTA", "TTGCA", 2 if multi else 1,
                    dp, nc, bcs, ccs, vaf, mcf, f("0.0" for _ in names), f("0.0" for _ in names), 2, 2, "0;10;1", "0;5;1", ".",
                    ctf, "DP|NC|CC|BC|BQ|BCf|BCr", info["Non-Cancer"], info["Cancer"]])
    return out


def write_table(path, seed=20260101, n=420):
    with open(path, "w") as o:
        o.write(HEADER)
        o.write("\t".join(COLUMNS) + "\n")
        for r in rows(seed, n):
            o.write("\t".join(str(x) for x in r) + "\n")
    return path


def write_fusions(path, barcodes, seed=7):
    """Fabricated CTAT-fusion table for CellTypeReannotation: '#FusionName', 'BC' (+ one ignored column),
    with repeated (fusion, barcode) pairs and barcodes that are absent from the genotype table."""
    rng = random.Random(seed)
    names = ["GENEA--GENEB", "GENEC--GENED", "GENEE--GENEF"]
    with open(path, "w") as o:
        o.write("#FusionName\tBC\tJunctionReadCount\n")
        picks = [b for i, b in enumerate(barcodes) if i % 5 == 0]
        for b in picks:
            for _ in range(rng.choice([1, 1, 2, 3])):
                o.write("%s\t%s\t%d\n" % (rng.choice(names), b, rng.randrange(1, 9)))
        o.write("GENEA--GENEB\tNOT_A_KNOWN_BARCODE\t2\n")
    return path
