#!/bin/bash
# usage: gpu_prof1.sh TAG [kernel-regex]  -- one --set full capture of one kernel of the full-scale step (after a plain run exited 0)
TAG=$1
K=${2:-pileup_count}
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --e2e-shards 1 --cli-scale 0"
timeout 900 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture exit $?"
