"""Measure max |p_gpu - p_scipy| of K2 by range of n (documentation aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import betabinom
from longsom_b200.engine import Engine
rng = np.random.default_rng(11)
a, b = 0.21356677091082193, 104.95163748636298
with Engine(0) as e:
    for lo, hi, m in ((1, 170, 20000), (170, 10000, 5000), (10000, 50000, 300), (50000, 200001, 120)):
        n = rng.integers(lo, hi, m).astype(np.int32)
        k = np.minimum((rng.random(m) ** 2 * np.minimum(n, 3000)).astype(np.int32) + 1, n).astype(np.int32)
        p = e.betabinom_sf(k, n, a, b)
        ref = betabinom.sf(k - 0.1, n, a, b)
        d = np.abs(p - ref)
        rel = d[ref > 1e-9] / ref[ref > 1e-9]
        print("n in [%d,%d): max abs %.3e  max rel(p>1e-9) %.3e  exact %d/%d  round4 mismatches %d  ms %.2f" % (
            lo, hi, d.max(), rel.max() if len(rel) else 0, int((p == ref).sum()), m,
            int((np.round(p, 4) != np.round(ref, 4)).sum()), e.last_stats["ms_count"]))
