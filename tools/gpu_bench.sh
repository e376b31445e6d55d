#!/bin/bash
# usage: gpu_bench.sh TAG [bench args]   -- one default bench run, JSON kept
TAG=$1; shift
cd "$GRAFT_REPO_ROOT"
timeout 1500 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"
tail -3 gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.4g ms_step %.3f k_ms %.3f frac %.4f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
print('stats', d['stats'])
print('e2e', d['e2e'])
print('secondary', json.dumps(d['secondary'])[:900])
print('c4', d.get('c4'))
print('cpu', d['cpu_baseline'])"
