#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tail -s 24 -c 1 -o gpurun_out/${TAG}_tail_prof -f python tools/tail_probe.py > gpurun_out/${TAG}_ncu.log 2>&1
( time timeout 600 python -m pytest tests -m gpu -q -x -k "next_scope or head or golden" ) > gpurun_out/${TAG}_pytest.log 2>&1
python - > gpurun_out/${TAG}_head.json 2> gpurun_out/${TAG}_head.err <<'PY'
import json, torch, sys
sys.path.insert(0, '.')
import bench
r = bench.head_sweep_leg(torch.device('cuda:0'), 'bf16')
print(json.dumps(r))
PY
tail -3 gpurun_out/${TAG}_ncu.log; tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_head.err
