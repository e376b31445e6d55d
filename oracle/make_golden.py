#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference scripts
(/root/reference/workflow/scripts/...) over oracle/shims on deterministic synthetic inputs.

Run in the build container only (the reference tree does not travel to the GPU box):
    python oracle/make_golden.py [case ...]
Outputs: tests/golden/<case>/*.gz  (+ manifest.json with the command lines used).
The reference ships no test vectors of its own; these files are what pins the
CPU oracle and the CUDA path to the reference's behaviour.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
REF = "/root/reference/workflow/scripts"
SHIMS = os.path.join(ROOT, "oracle", "shims")


def run_ref(script, args, log, produces=None):
    if produces and os.path.exists(produces) and os.environ.get("LS_GOLDEN_RESUME"):
        log.append("(resumed) " + script)
        return ""
    env = dict(os.environ)
    env["PYTHONPATH"] = SHIMS + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, os.path.join(REF, script)] + [str(a) for a in args]
    log.append(" ".join(cmd[1:]))
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference script failed: %s\n%s" % (" ".join(cmd), r.stdout[-3000:]))
    return r.stdout


def candidates_from_step2(step2_path, out_path):
    """Deterministic candidate list for the genotype scripts: every PASS row plus every tenth other row
    (a wider set than step3's PASS rows, which are too few here to exercise the genotype kernels; same column layout)."""
    k = 0
    with open(step2_path) as f, open(out_path, "w") as o:
        for line in f:
            if line.startswith("#"):
                o.write(line)
                continue
            cols = line.rstrip("\n").split("\t")
            if cols[5] == "PASS" or k % 10 == 0:
                o.write(line)
            k += 1


# (name suffix, deltaVAF, deltaMCF, min_ac_reads, min_ac_cells, clust_dist): the workflow's LongSom settings
# (config.yaml:81-83 with the 10 kb cluster distance) and a loose set that leaves PASS / clustered rows
STEP3_PARAMS = (("step3", 0.05, 0.3, 3, 2, 10000), ("step3_loose", 0.05, 0.05, 1, 1, 50))


def step3_stage(step2_path, out, log, param_sets=STEP3_PARAMS):
    """BaseCellCalling.step3 (8f-1) on a step2 table -> {golden name: path}."""
    res = {}
    for name, dvaf, dmcf, mr, mc, cd in param_sets:
        pre = os.path.join(out, name)
        run_ref("SNVCalling/BaseCellCalling.step3.py", ["--infile", step2_path, "--outfile", pre, "--deltaVAF", dvaf,
                "--deltaMCF", dmcf, "--min_ac_reads", mr, "--min_ac_cells", mc, "--clust_dist", cd], log)
        res[name + ".tsv"] = pre + ".calling.step3.tsv"
        res[name + ".unfiltered.tsv"] = pre + ".calling.step3.unfiltered.tsv"
    return res


# (name, min_dp, deltaVAF, deltaMCF, clust_dist): script defaults (= config.yaml CellTypeReannotation) and a loose set
HCCV_PARAMS = (("hccv_variants", 20, 0.1, 0.4, 10000), ("hccv_variants_loose", 5, 0.05, 0.05, 30))


def hccv_variants_stage(step2_path, out, log, param_sets=HCCV_PARAMS):
    """HighConfidenceCancerVariants (8f-2) on a step2 table -> {golden name: path}.  A parameter set that leaves
    no variant makes the reference fail inside pandas after writing its three files; that case is skipped here."""
    res = {}
    for name, min_dp, dvaf, dmcf, cd in param_sets:
        pre = os.path.join(out, name)
        try:
            run_ref("CellTypeReannotation/HighConfidenceCancerVariants.py", ["--SNVs", step2_path, "--outfile", pre,
                    "--min_dp", min_dp, "--deltaVAF", dvaf, "--deltaMCF", dmcf, "--clust_dist", cd], log)
        except RuntimeError as e:
            log.append("(reference failed, no golden) " + name + ": " + str(e).strip().splitlines()[-1][:120])
            continue
        for suf in ("", "2", "3"):
            res[name + ".tsv" + suf] = pre + ".HCCV.tsv" + suf
    return res


REANNOT_PARAMS = (("reannot", 3, 0.2), ("reannot_loose", 1, 0.02), ("reannot_mid", 5, 0.03))   # (name, min_variants, min_frac)


def reannotation_stage(hccv_genotypes, meta, out, log):
    """CellTypeReannotation (8f-2) on the HCCV genotype table + a fabricated fusion table -> {golden name: path}."""
    import step3_inputs
    barcodes = [l.split("\t")[0] for l in open(meta)][1:]
    fusions = step3_inputs.write_fusions(os.path.join(out, "reannot_fusions.tsv"), barcodes)
    res = {"reannot_fusions.tsv": fusions}
    for name, mv, mf in REANNOT_PARAMS:
        dst = os.path.join(out, name + ".tsv")
        run_ref("CellTypeReannotation/CellTypeReannotation.py", ["--SNVs", hccv_genotypes, "--fusions", fusions, "--outfile",
                dst, "--meta", meta, "--min_variants", mv, "--min_frac", mf], log)
        res[name + ".tsv"] = dst
    return res


# (name, extra CLI arguments): defaults of the workflow (PreProcessing.smk passes --min_MQ and --n_trim) and a set
# with every filter on
SPLIT_PARAMS = (("split", ["--min_MQ", 60, "--n_trim", 5]),
                ("split_all", ["--min_MQ", 30, "--n_trim", 0, "--max_nM", 5, "--max_NH", 1]))


def splitbam_stage(case, work, log):
    """SplitBamCellTypes (8f-3) over the shim's BAM writer -> {golden name: path}: per output BAM a text dump of
    its records (tests/support/pipeline_inputs.dump_bam_records) and the report without its run-time column."""
    import pipeline_inputs as pi
    bam, meta = pi.write_split_input(case, work)
    res = {}
    for name, extra in SPLIT_PARAMS:
        outdir = os.path.join(work, name)
        os.makedirs(outdir, exist_ok=True)
        run_ref("PreProcessing/SplitBamCellTypes.py", ["--bam", bam, "--meta", meta, "--id", "s", "--outdir", outdir] + extra, log)
        for fn in sorted(os.listdir(outdir)):
            if fn.endswith(".bam"):
                dump = os.path.join(outdir, fn + ".records.txt")
                with open(dump, "w") as o:
                    o.write("\n".join(pi.dump_bam_records(os.path.join(outdir, fn))) + "\n")
                res["%s.%s.records.txt" % (name, fn[:-4])] = dump
                assert os.path.exists(os.path.join(outdir, fn + ".bai"))
        rep = os.path.join(outdir, "s.report.txt")
        rows = [l.rstrip("\n").split("\t") for l in open(rep)]
        keep = [i for i, k in enumerate(rows[0]) if k != "Total_time"]
        with open(rep + ".notime", "w") as o:
            for r in rows:
                o.write("\t".join(r[i] for i in keep) + "\n")
        res[name + ".report.txt"] = rep + ".notime"
    return res


def write_golden(gdir, res):
    os.makedirs(gdir, exist_ok=True)
    sizes = {}
    for name, path in res.items():
        data = open(path, "rb").read()
        with gzip.GzipFile(os.path.join(gdir, name + ".gz"), "wb", mtime=0) as g:
            g.write(data)
        sizes[name] = len(data)
    return sizes


def step3_only(cases):
    """`make_golden.py --step3 [case ...]`: (re)generate only the goldens of the host-only stages after the path
    (step3, HighConfidenceCancerVariants, CellTypeReannotation) from the committed step2 / HCCV-genotype goldens
    (g1, g2) and from the fabricated branch-coverage table (s3); other golden files are untouched."""
    import step3_inputs
    for case in cases:
        work = tempfile.mkdtemp(prefix="ls_golden_s3_")
        log, gdir = [], os.path.join(ROOT, "tests", "golden", case)
        if case == "s3":
            table = step3_inputs.write_table(os.path.join(work, "step2_fabricated.tsv"))
            res = {"step2_fabricated.tsv": table}
            res.update(step3_stage(table, work, log, (("step3", 0.2, 0.25, 3, 2, 1000),)))
            res.update(hccv_variants_stage(table, work, log))
            manifest = {"case": case, "generated_by": "oracle/make_golden.py --step3 (tests/support/step3_inputs.py -> "
                        "reference BaseCellCalling.step3.py)"}
        else:
            table = os.path.join(work, "step2.tsv")
            with gzip.open(os.path.join(gdir, "step2.tsv.gz"), "rb") as f, open(table, "wb") as o:
                o.write(f.read())
            res = step3_stage(table, work, log)
            res.update(hccv_variants_stage(table, work, log))
            import pipeline_inputs as pi
            paths, _ = pi.write_inputs(case, work)
            geno = os.path.join(work, "hccv.tsv")
            with gzip.open(os.path.join(gdir, "hccv.tsv.gz"), "rb") as f, open(geno, "wb") as o:
                o.write(f.read())
            res.update(reannotation_stage(geno, paths["meta"], work, log))
            res.update(splitbam_stage(case, work, log))
            manifest = json.load(open(os.path.join(gdir, "manifest.json")))
        sizes = write_golden(gdir, res)
        manifest.setdefault("bytes", {}).update(sizes)
        manifest["reference_commands"] = [c for c in manifest.get("reference_commands", []) if not any(w in c for w in ("step3", "HighConfidence", "hccv_variants", "CellTypeReannotation.py", "SplitBam"))] + log
        json.dump(manifest, open(os.path.join(gdir, "manifest.json"), "w"), indent=1, default=str)
        print(case, "->", gdir, sizes)
        shutil.rmtree(work, ignore_errors=True)


def bed_stage(case, work, log):
    """BaseCellCounter --bed / --bed_out (MakeWindows over the pybedtools stand-in) -> {golden name: path}."""
    import pipeline_inputs as pi
    p, d = pi.write_inputs(case, work)
    bed, bed_out = pi.write_beds(case, work, d)
    out = os.path.join(work, "out", "counts_bed")
    os.makedirs(out, exist_ok=True)
    res = {}
    for name, extra in (("bed", ["--bed", bed, "--chrom", "all"]),
                        ("bed_out", ["--bed", bed, "--bed_out", bed_out, "--chrom", d.contig_names[0]]),
                        ("bedout_only", ["--bed_out", bed_out, "--bin", 20000, "--chrom", "all"])):
        run_ref("SNVCalling/BaseCellCounter.py", ["--bam", p["full"], "--ref", p["ref"], "--out_folder", out, "--id", "full." + name,
                "--min_bq", 20, "--min_mq", 60, "--min_dp", 3, "--min_cc", 2, "--nprocs", 1, "--tmp_dir",
                os.path.join(work, "tmp_" + name)] + extra, log, produces=os.path.join(out, "full.%s.tsv" % name))
        res["counts.full_%s.tsv" % name] = os.path.join(out, "full.%s.tsv" % name)
    return res


def bed_only(cases):
    """`make_golden.py --bed [case ...]`: add the --bed / --bed_out BaseCellCounter goldens to existing cases."""
    for case in cases:
        work = tempfile.mkdtemp(prefix="ls_golden_bed_")
        log, gdir = [], os.path.join(ROOT, "tests", "golden", case)
        res = bed_stage(case, work, log)
        sizes = write_golden(gdir, res)
        manifest = json.load(open(os.path.join(gdir, "manifest.json")))
        manifest.setdefault("bytes", {}).update(sizes)
        manifest["reference_commands"] = [c for c in manifest.get("reference_commands", []) if "--bed" not in c] + log
        json.dump(manifest, open(os.path.join(gdir, "manifest.json"), "w"), indent=1, default=str)
        print(case, "->", gdir, sizes)
        shutil.rmtree(work, ignore_errors=True)


def reference_pipeline(case, work, log):
    import pipeline_inputs as pi
    p, d = pi.write_inputs(case, work)
    out = os.path.join(work, "out")
    os.makedirs(os.path.join(out, "counts"), exist_ok=True)
    res = {}
    core_only = case in pi.CORE_ONLY
    for name, bam in (("Cancer", p["cancer"]), ("Non-Cancer", p["normal"])):
        run_ref("SNVCalling/BaseCellCounter.py", ["--bam", bam, "--ref", p["ref"], "--chrom", "all", "--out_folder",
                os.path.join(out, "counts"), "--min_bq", 20, "--min_mq", 60, "--nprocs", 1, "--tmp_dir",
                os.path.join(work, "tmp_" + name)], log, produces=os.path.join(out, "counts", "s.%s.tsv" % name))
        res["counts.%s.tsv" % name] = os.path.join(out, "counts", "s.%s.tsv" % name)
    if not core_only:
        # a second parameterisation on the un-split BAM: min_ac > 0 exercises the AC pre-gate (Q4)
        os.makedirs(os.path.join(out, "counts_ac"), exist_ok=True)
        run_ref("SNVCalling/BaseCellCounter.py", ["--bam", p["full"], "--ref", p["ref"], "--chrom", d.contig_names[0],
                "--out_folder", os.path.join(out, "counts_ac"), "--id", "full.ac", "--min_bq", 30, "--min_mq", 0, "--min_ac", 2,
                "--min_dp", 3, "--min_cc", 2, "--bin", 30000, "--nprocs", 1, "--tmp_dir", os.path.join(work, "tmp_ac")], log,
                produces=os.path.join(out, "counts_ac", "full.ac.tsv"))
        res["counts.full_ac.tsv"] = os.path.join(out, "counts_ac", "full.ac.tsv")
    merged = os.path.join(out, "merged.tsv")
    run_ref("SNVCalling/MergeBaseCellCounts.py", ["--tsv_folder", os.path.join(out, "counts"), "--outfile", merged], log)
    res["merged.tsv"] = merged
    run_ref("SNVCalling/BaseCellCalling.step1.py", ["--infile", merged, "--outfile", os.path.join(out, "s"), "--ref", p["ref"],
            "--min_cell_types", 2, "--min_ac_reads", 3, "--min_ac_cells", 2, "--alpha1", pi.ALPHA1, "--beta1", pi.BETA1,
            "--alpha2", pi.ALPHA2, "--beta2", pi.BETA2], log)
    res["step1.tsv"] = os.path.join(out, "s.calling.step1.tsv")
    step2 = os.path.join(out, "s.calling.step2.tsv")
    run_ref("SNVCalling/BaseCellCalling.step2.py", ["--infile", res["step1.tsv"], "--outfile", os.path.join(out, "s"), "--editing", p["editing"],
            "--pon_SR", p["pon_sr"], "--pon_LR", p["pon_lr"], "--gnomAD_db", p["gnomad"], "--gnomAD_max", 0.01,
            "--min_distance", 0], log)  # the workflow passes 0 (SNVCalling.smk)
    res["step2.tsv"] = step2
    cand = os.path.join(out, "candidates.tsv")
    if core_only:
        candidates_from_step2(step2, cand)
        res["candidates.tsv"] = cand
        pre = os.path.join(out, "geno_All")
        run_ref("CellClustering/SingleCellGenotype.py", ["--bam", p["full"], "--infile", cand, "--ref", p["ref"], "--meta",
                p["meta"], "--fusions", "--outfile", pre, "--alt_flag", "All", "--nprocs", 1, "--min_mq", 60, "--pvalue", 0.01,
                "--alpha2", pi.ALPHA2, "--beta2", pi.BETA2, "--chrM_contaminant", "True", "--tmp_dir",
                os.path.join(work, "tmp_gAll")], log)
        for suf in ("SingleCellGenotype", "DpMatrix", "AltMatrix", "VAFMatrix", "BinaryMatrix"):
            res["geno_All.%s.tsv" % suf] = "%s.%s.tsv" % (pre, suf)
        return res
    # the shipped config points --editing at a .gz file: the filter is silently off (Q9)
    step2gz = os.path.join(out, "s.gz.calling.step2.tsv")
    run_ref("SNVCalling/BaseCellCalling.step2.py", ["--infile", res["step1.tsv"], "--outfile", os.path.join(out, "s.gz"), "--editing",
            p["editing_gz"], "--pon_SR", p["pon_sr"], "--pon_LR", "--gnomAD_db", p["gnomad"], "--gnomAD_max", 0.01,
            "--min_distance", 5], log)  # script default: exercises the 3-row 'Clustered' logic
    res["step2_gz.tsv"] = step2gz
    res.update(step3_stage(step2, out, log))
    res.update(hccv_variants_stage(step2, out, log))
    candidates_from_step2(step2, cand)
    res["candidates.tsv"] = cand
    for flag in ("All", "Alt"):
        pre = os.path.join(out, "geno_" + flag)
        run_ref("CellClustering/SingleCellGenotype.py", ["--bam", p["full"], "--infile", cand, "--ref", p["ref"], "--meta",
                p["meta"], "--fusions", "--outfile", pre, "--alt_flag", flag, "--nprocs", 1, "--min_mq", 60, "--pvalue", 0.01,
                "--alpha2", pi.ALPHA2, "--beta2", pi.BETA2, "--chrM_contaminant", "True", "--tmp_dir",
                os.path.join(work, "tmp_g" + flag)], log)
        for suf in ("SingleCellGenotype", "DpMatrix", "AltMatrix", "VAFMatrix", "BinaryMatrix"):
            res["geno_%s.%s.tsv" % (flag, suf)] = "%s.%s.tsv" % (pre, suf)
    hccv = os.path.join(out, "hccv.tsv")
    run_ref("CellTypeReannotation/HCCVSingleCellGenotype.py", ["--bam", p["full"], "--infile", cand, "--ref", p["ref"], "--meta",
            p["meta"], "--outfile", hccv, "--alt_flag", "All", "--nprocs", 1, "--min_mq", 60, "--pvalue", 0.01,
            "--chrM_contaminant", "True", "--tmp_dir", os.path.join(work, "tmp_h")], log)
    res["hccv.tsv"] = hccv
    res.update(reannotation_stage(hccv, p["meta"], out, log))
    res.update(splitbam_stage(case, work, log))
    return res


def main():
    import pipeline_inputs as pi
    if sys.argv[1:2] == ["--step3"]:
        return step3_only(sys.argv[2:] or [c for c in pi.CASES if c not in pi.CORE_ONLY] + ["s3"])
    if sys.argv[1:2] == ["--bed"]:
        return bed_only(sys.argv[2:] or ["g1"])
    cases = sys.argv[1:] or list(pi.CASES)
    for case in cases:
        work = os.path.join(tempfile.gettempdir(), "ls_golden_work_%s" % case)
        os.makedirs(work, exist_ok=True)
        log = []
        res = reference_pipeline(case, work, log)
        gdir = os.path.join(ROOT, "tests", "golden", case)
        sizes = write_golden(gdir, res)
        json.dump({"case": case, "synth": pi.CASES[case], "reference_commands": log, "bytes": sizes,
                   "generated_by": "oracle/make_golden.py (reference scripts from /root/reference over oracle/shims)"},
                  open(os.path.join(gdir, "manifest.json"), "w"), indent=1, default=str)
        print(case, "->", gdir, {k: v for k, v in sizes.items()})
        if not os.environ.get("LS_GOLDEN_KEEP"):
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
