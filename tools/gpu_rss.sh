#!/bin/bash
# usage: gpu_rss.sh SCALE -- one streaming BaseCellCounter run on a synthetic BAM with the resident-set breakdown
cd "$GRAFT_REPO_ROOT"
python - "$1" <<'PY'
import os, sys, time, tempfile, subprocess
sys.path.insert(0, os.getcwd())
from longsom_b200 import synth, bamio
scale = float(sys.argv[1])
d = synth.generate(**synth.config("C2", scale=scale)); b = d.batch
tmp = tempfile.mkdtemp()
bamio.write_fasta(tmp + "/ref.fa", d.contig_names, [d.contig_seq(i) for i in range(len(d.contig_lens))])
names = [synth.barcode_of(c) + "-1" for c in range(d.n_cells + d.n_extra_cells)]
bamio.write_bam(tmp + "/x.bam", d.contig_names, d.contig_lens, b, lambda i: None if b.cell[i] < 0 else names[b.cell[i]])
del d, b
for mb in ("256", "64"):
    e = dict(os.environ, LS_STREAM_TIMING="2", LONGSOM_CHUNK_MB=mb, LS_BAM_TIMING="1")
    t0 = time.time()
    r = subprocess.run([sys.executable, "workflow/scripts/SNVCalling/BaseCellCounter.py", "--bam", tmp + "/x.bam", "--ref", tmp + "/ref.fa",
                        "--chrom", "all", "--out_folder", tmp, "--id", "x" + mb, "--min_bq", "20", "--min_mq", "60"], env=e, capture_output=True, text=True)
    print("chunk MB", mb, "seconds %.2f" % (time.time() - t0), "rc", r.returncode)
    print(r.stderr[-3000:])
PY
