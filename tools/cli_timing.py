"""Wall-clock breakdown of the BaseCellCounter drop-in on a synthetic BAM (decode / GPU / TSV)."""
import os, sys, time, tempfile, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from longsom_b200 import synth, bamio
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.04
d = synth.generate(**synth.config("C2", scale=scale))
tmp = tempfile.mkdtemp()
b = d.batch
t = time.time()
bamio.write_fasta(tmp + "/ref.fa", d.contig_names, [d.contig_seq(i) for i in range(len(d.contig_lens))])
names = [synth.barcode_of(c) + "-1" for c in range(d.n_cells + d.n_extra_cells)]
bamio.write_bam(tmp + "/x.bam", d.contig_names, d.contig_lens, b, lambda i: None if b.cell[i] < 0 else names[b.cell[i]])
print("reads %d, aligned bases %d, BAM %.1f MB (written in %.1f s)" % (b.n_reads, b.aligned_bases(), os.path.getsize(tmp + "/x.bam") / 1e6, time.time() - t))
from longsom_b200.pipeline import load_bam_for_counting, prune_and_sort_windows, read_ends, count_sites, write_counter_tsv, prewarm
from longsom_b200.windows import make_windows
from longsom_b200.engine import CountParams
t0 = time.time(); prewarm([0]); fa = bamio.Fasta(tmp + "/ref.fa"); bd, batch, _ = load_bam_for_counting(tmp + "/x.bam"); t1 = time.time()
named = make_windows(fa.references, fa.lengths, "all", 50000)
iv = prune_and_sort_windows(named, bd.contig_names, batch, read_ends(batch))
seqs = {t_: fa.contig(bd.contig_names[t_]) for t_ in sorted({w[0] for w in iv})}; t2 = time.time()
stats = []
sites = count_sites(batch, iv, seqs, CountParams(min_bq=20, min_mq=60), [0], stats); t3 = time.time()
write_counter_tsv(tmp + "/out.tsv", "x", sites, bd.contig_names); t4 = time.time()
print("decode BAM %.2f s | windows+ref %.2f s | GPU upload+run+fetch %.2f s (device %.1f ms) | TSV %d sites %.2f s | total %.2f s" % (
    t1 - t0, t2 - t1, t3 - t2, stats[0]["ms_total"], sites.n_sites, t4 - t3, t4 - t0))
print("end-to-end %.3g aligned bases/s from BAM on disk to TSV on disk" % (b.aligned_bases() / (t4 - t0)))
# the same GPU section again (CUDA context and library already warm), one call and pipelined window shards
for shards in ("1", "8"):
    os.environ["LONGSOM_SHARDS_PER_GPU"] = shards
    ta = time.time()
    s2 = count_sites(batch, iv, seqs, CountParams(min_bq=20, min_mq=60), [0], [])
    print("warm GPU section, LONGSOM_SHARDS_PER_GPU=%s: %.2f s (%d sites)" % (shards, time.time() - ta, s2.n_sites))
    assert s2.n_sites == sites.n_sites and np.array_equal(s2.counts, sites.counts)
