import sys; sys.path.insert(0,'.')
from longsom_b200 import synth
from longsom_b200.batch import Windows, make_windows
from longsom_b200.engine import CountParams, Engine
d = synth.generate(seed=8, contig_lens=[500000], n_genes=8, n_reads=2500, n_cells=40, mean_len=6000.0)
w = Windows.from_intervals(make_windows(d.contig_lens, 50000), d.contig_seqs())
with Engine(0) as e:
    e.upload(d.batch, w)
    for k in range(2):
        n = e.run(CountParams(min_bq=20, min_mq=60)); print('run', k, 'sites', n, 'launches', e.last_stats['count_launches'], 'segments', e.last_stats['n_segments'], 'reads', d.batch.n_reads)
