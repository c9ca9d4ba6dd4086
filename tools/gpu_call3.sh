#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tail -s 2 -c 2 -o gpurun_out/${TAG}_tail_prof -f python tools/tail_probe.py > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/ | tail -5
tail -5 gpurun_out/${TAG}_ncu.log
