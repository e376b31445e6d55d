#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_pileup_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/d_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/d_tests.log
tail -5 gpurun_out/d_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-shards 1 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/d_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['stats'])"
CMD="python bench.py --scale 0.2 --steps 1 --warmup 1 --no-cpu --e2e-shards 1"
timeout 600 $CMD > gpurun_out/d_plain.json 2> gpurun_out/d_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pileup_count -s 1 -c 1 -o gpurun_out/prof_k1_v7 -f $CMD > gpurun_out/d_ncu.log 2>&1
echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/d_launches.csv $CMD > gpurun_out/d_ncu2.log 2>&1
