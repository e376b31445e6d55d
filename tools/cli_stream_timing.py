"""Disk-to-disk wall clock of the drop-in BaseCellCounter on a synthetic C2 BAM: streaming path (default) against the
whole-file path (LONGSOM_STREAM=0), byte-compared, with the peak resident set of each run.
usage: cli_stream_timing.py [scale] [chunk_mb]"""
import os, sys, time, tempfile, subprocess, resource, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from longsom_b200 import synth, bamio
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
d = synth.generate(**synth.config("C2", scale=scale))
tmp = tempfile.mkdtemp()
b = d.batch
bamio.write_fasta(tmp + "/ref.fa", d.contig_names, [d.contig_seq(i) for i in range(len(d.contig_lens))])
names = [synth.barcode_of(c) + "-1" for c in range(d.n_cells + d.n_extra_cells)]
t = time.time()
bamio.write_bam(tmp + "/x.bam", d.contig_names, d.contig_lens, b, lambda i: None if b.cell[i] < 0 else names[b.cell[i]])
aligned = b.aligned_bases()
print("reads %d, aligned bases %d, BAM %.1f MB (written in %.1f s), %d host cores" % (
    b.n_reads, aligned, os.path.getsize(tmp + "/x.bam") / 1e6, time.time() - t, os.cpu_count()), flush=True)
del d, b
script = os.path.join(ROOT, "workflow", "scripts", "SNVCalling", "BaseCellCounter.py")
res = {}
for mode, env in (("stream", {}), ("whole", {"LONGSOM_STREAM": "0"})):  # streaming first: the smaller footprint
    e = dict(os.environ, **env)
    if len(sys.argv) > 2 and mode == "stream":
        e["LONGSOM_CHUNK_MB"] = sys.argv[2]
    out = tmp + "/out_" + mode
    os.makedirs(out, exist_ok=True)
    best = None
    e["LS_STREAM_TIMING"] = "1"
    for rep in range(3):
        t0 = time.time()
        rss0 = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
        r = subprocess.run([sys.executable, script, "--bam", tmp + "/x.bam", "--ref", tmp + "/ref.fa", "--chrom", "all",
                            "--out_folder", out, "--id", "x", "--min_bq", "20", "--min_mq", "60", "--tmp_dir", out + "/tmp"],
                           env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        dt = time.time() - t0
        assert r.returncode == 0, r.stderr[-2000:]
        rss_gb = None
        for l in r.stderr.splitlines():
            if l.startswith("[stream_count]"):
                print("   ", l, flush=True)
            if l.startswith("[counter] peak RSS"):   # VmHWM of the child (ru_maxrss would carry this parent's RSS over exec)
                rss_gb = float(l.split()[3])
        best = dt if best is None or dt < best else best
    res[mode] = dict(seconds=best, bases_per_s=aligned / best, peak_rss_gb=rss_gb)
    print("%-6s %.2f s  %.3g aligned bases/s disk to disk, peak RSS %.2f GB" % (mode, best, aligned / best, rss_gb or -1), flush=True)
a = open(tmp + "/out_stream/x.tsv").read().split("\n", 1)[1]
w = open(tmp + "/out_whole/x.tsv").read().split("\n", 1)[1]
assert a == w, "streaming and whole-file tables differ"
print("tables identical (%d bytes)" % len(a))
print(json.dumps({"scale": scale, "aligned_bases": aligned, **res}))
