#!/bin/bash
# usage: gpu_tests.sh TAG [pytest args]   -- run (a subset of) the GPU suite and keep the log
TAG=$1; shift
cd "$GRAFT_REPO_ROOT"
nproc
timeout 2400 python -m pytest "$@" -x -q --durations=15 > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -30 gpurun_out/${TAG}_tests.log
