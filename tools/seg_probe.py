#!/usr/bin/env python
"""Stage timings of ls_pileup_run on the C2 batch for a few parameter sets (segment builder with / without the
depth-cap bookkeeping, etc.).  Usage: python tools/seg_probe.py [scale]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from longsom_b200.engine import CountParams, Engine  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
wl = bench.build_workload(scale, 0, 1)
eng = Engine(0)
eng.upload(wl["batch"], wl["windows"])
for name, over in (("default", {}), ("max_depth=0 (no wcount / rend)", {"max_depth": 0}), ("min_ac=2 (uncounted reads emitted)", {"min_ac": 2})):
    p = dict(bench.PARAMS)
    p.update(over)
    prm = CountParams(**p)
    for _ in range(3):
        eng.run(prm)
    acc = {}
    for _ in range(5):
        eng.run(prm)
        for k in ("ms_segments", "ms_sort", "ms_count", "ms_total"):
            acc[k] = acc.get(k, 0.0) + eng.last_stats.get(k, 0.0) / 5
    print(name, {k: round(v, 3) for k, v in acc.items()}, flush=True)
