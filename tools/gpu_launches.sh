#!/bin/bash
# usage: gpu_launches.sh TAG  -- ncu launch list (gpu__time_duration per launch) of the full-scale step, after a plain run exited 0
TAG=$1
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --e2e-shards 1 --cli-scale 0"
timeout 900 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
