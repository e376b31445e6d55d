#!/bin/bash
TAG=$1
cd "$GRAFT_REPO_ROOT"
python - <<'PY' > gpurun_out/${TAG}_rss.log 2>&1
import resource, sys
sys.path.insert(0, ".")
r0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e3
from longsom_b200.engine import Engine
e = Engine(0)
r1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e3
print("RSS MB: python+numpy %.0f, after ls_ctx_create %.0f" % (r0, r1))
PY
cat gpurun_out/${TAG}_rss.log
LS_STREAM_TIMING=1 timeout 1500 python tools/cli_stream_timing.py 0.2 > gpurun_out/${TAG}_cli02.log 2>&1; tail -5 gpurun_out/${TAG}_cli02.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench exit $?
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.4g ms_step %.3f k_ms %.3f frac %.4f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
print('e2e_cli', d['e2e_cli'])"
