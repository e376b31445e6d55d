#!/bin/bash
# usage: gpu_ab.sh TAG "ENV1" "ENV2" ...   -- K1 parity tests once, then the C2 bench once per environment setting
TAG=$1; shift
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_pileup_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
n=0
for E in "$@"; do
  n=$((n+1))
  env $E timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-shards 1 > gpurun_out/${TAG}_bench$n.json 2> gpurun_out/${TAG}_bench$n.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench$n.json'))
print('[$E] value %.4g ms_step %.3f k_ms %.3f frac %.4f seg %.2f sort %.2f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['stats']['ms_segments'], d['stats']['ms_sort']))"
done
