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
    (the pipeline would use step3 output / HCCV, which are out of scope; same column layout)."""
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


def reference_pipeline(case, work, log):
    import pipeline_inputs as pi
    p, d = pi.write_inputs(case, work)
    out = os.path.join(work, "out")
    os.makedirs(os.path.join(out, "counts"), exist_ok=True)
    res = {}
    for name, bam in (("Cancer", p["cancer"]), ("Non-Cancer", p["normal"])):
        run_ref("SNVCalling/BaseCellCounter.py", ["--bam", bam, "--ref", p["ref"], "--chrom", "all", "--out_folder",
                os.path.join(out, "counts"), "--min_bq", 20, "--min_mq", 60, "--nprocs", 1, "--tmp_dir",
                os.path.join(work, "tmp_" + name)], log, produces=os.path.join(out, "counts", "s.%s.tsv" % name))
        res["counts.%s.tsv" % name] = os.path.join(out, "counts", "s.%s.tsv" % name)
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
    # the shipped config points --editing at a .gz file: the filter is silently off (Q9)
    step2gz = os.path.join(out, "s.gz.calling.step2.tsv")
    run_ref("SNVCalling/BaseCellCalling.step2.py", ["--infile", res["step1.tsv"], "--outfile", os.path.join(out, "s.gz"), "--editing",
            p["editing_gz"], "--pon_SR", p["pon_sr"], "--pon_LR", "--gnomAD_db", p["gnomad"], "--gnomAD_max", 0.01,
            "--min_distance", 5], log)  # script default: exercises the 3-row 'Clustered' logic
    res["step2_gz.tsv"] = step2gz
    cand = os.path.join(out, "candidates.tsv")
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
    return res


def main():
    import pipeline_inputs as pi
    cases = sys.argv[1:] or list(pi.CASES)
    for case in cases:
        work = os.path.join(tempfile.gettempdir(), "ls_golden_work_%s" % case)
        os.makedirs(work, exist_ok=True)
        log = []
        res = reference_pipeline(case, work, log)
        gdir = os.path.join(ROOT, "tests", "golden", case)
        os.makedirs(gdir, exist_ok=True)
        sizes = {}
        for name, path in res.items():
            data = open(path, "rb").read()
            with gzip.GzipFile(os.path.join(gdir, name + ".gz"), "wb", mtime=0) as g:
                g.write(data)
            sizes[name] = len(data)
        json.dump({"case": case, "synth": pi.CASES[case], "reference_commands": log, "bytes": sizes,
                   "generated_by": "oracle/make_golden.py (reference scripts from /root/reference over oracle/shims)"},
                  open(os.path.join(gdir, "manifest.json"), "w"), indent=1, default=str)
        print(case, "->", gdir, {k: v for k, v in sizes.items()})
        if not os.environ.get("LS_GOLDEN_KEEP"):
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
