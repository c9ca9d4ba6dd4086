#!/bin/bash
# block_tail2 bring-up: unit parity, stand-alone timing against the first-generation kernel
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_block_tail.py -x -q ) > gpurun_out/t2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/t2_pytest.log
( timeout 200 python tools/tail_probe.py ) > gpurun_out/t2_probe.txt 2>&1
( TANTE_TAIL_STREAMS=1 timeout 200 python tools/tail_probe.py ) > gpurun_out/t2_probe_v1.txt 2>&1
tail -15 gpurun_out/t2_pytest.log; cat gpurun_out/t2_probe.txt gpurun_out/t2_probe_v1.txt
