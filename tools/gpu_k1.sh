#!/bin/bash
# usage: gpu_k1.sh TAG [ncu]   -- K1 parity tests, C2 bench (device-resident), optional ncu capture at scale 0.2
TAG=$1
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_pileup_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-shards 1 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.4g ms_step %.3f k_ms %.3f frac %.4f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']), d['stats'])"
if [ "$2" == "ncu" ]; then
CMD="python bench.py --scale 0.2 --steps 1 --warmup 1 --no-cpu --e2e-shards 1"
timeout 600 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pileup_count|expand_kernel" -s 2 -c 2 -o gpurun_out/prof_k1_${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
fi
