#!/bin/bash
# round-2 GPU call A: K1 parity + first timing of the lane-per-unit kernel
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_pileup_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-shards 1 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?"
cat gpurun_out/a_bench.json
