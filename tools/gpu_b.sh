#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 ./tools/ubench_atoms2 > gpurun_out/b_ubench.log 2>&1
cat gpurun_out/b_ubench.log
CMD="python bench.py --scale 0.2 --steps 1 --warmup 1 --no-cpu --e2e-shards 1"
timeout 600 $CMD > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pileup_count -s 1 -c 1 -o gpurun_out/prof_k1_v5 -f $CMD > gpurun_out/b_ncu.log 2>&1
echo "ncu exit $?"
tail -3 gpurun_out/b_ncu.log
