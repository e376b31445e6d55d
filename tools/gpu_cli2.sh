#!/bin/bash
TAG=$1
cd "$GRAFT_REPO_ROOT"
timeout 1500 python tools/cli_stream_timing.py 0.2 > gpurun_out/${TAG}_cli02.log 2>&1; tail -6 gpurun_out/${TAG}_cli02.log
timeout 2400 python tools/cli_stream_timing.py 1.0 > gpurun_out/${TAG}_cli10.log 2>&1; tail -6 gpurun_out/${TAG}_cli10.log
