#!/bin/bash
# usage: gpu_launches2.sh TAG  -- ncu launch list of a bench run WITH the genotyping block (after a plain run exited 0)
TAG=$1
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-shards 1 --cli-scale 0"
timeout 900 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
