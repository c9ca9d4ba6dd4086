#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-t2}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tail2 -s 24 -c 1 -o gpurun_out/${TAG}_tail2_prof -f python tools/tail_probe.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -5 gpurun_out/${TAG}_ncu.log
