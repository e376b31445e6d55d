#!/bin/bash
# usage: gpu_prof.sh TAG   -- full-scale ncu evidence for the count kernel: launch list + one --set full capture
TAG=$1
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --e2e-shards 1"
timeout 900 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
timeout 900 $CMD > /dev/null 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pileup_count -s 2 -c 1 -o gpurun_out/prof_k1_${TAG} -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture exit $?"
