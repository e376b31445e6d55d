#!/bin/bash
TAG=$1
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_cli_gpu.py -x -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
LONGSOM_CHUNK_MB=0.3 timeout 900 python -m pytest tests/test_cli_gpu.py -x -q -k "base_cell_counter or multi_device" > gpurun_out/${TAG}_tests_small.log 2>&1
echo "small-chunk tests exit $?" | tee -a gpurun_out/${TAG}_tests_small.log
tail -3 gpurun_out/${TAG}_tests_small.log
timeout 1500 python tools/cli_stream_timing.py 0.2 > gpurun_out/${TAG}_cli02.log 2>&1; tail -6 gpurun_out/${TAG}_cli02.log
timeout 2400 python tools/cli_stream_timing.py 1.0 > gpurun_out/${TAG}_cli10.log 2>&1; tail -6 gpurun_out/${TAG}_cli10.log
