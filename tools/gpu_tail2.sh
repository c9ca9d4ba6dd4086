#!/bin/bash
# block-tail bring-up: unit parity (pairs where they apply / forced everywhere / off), stand-alone timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_block_tail.py -x -q ) > gpurun_out/t2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/t2_pytest.log
( TANTE_TAIL_2CTA=2 timeout 600 python -m pytest tests/test_gpu_block_tail.py -x -q ) >> gpurun_out/t2_pytest.log 2>&1
echo "pytest (pairs forced) exit $?" >> gpurun_out/t2_pytest.log
( timeout 200 python tools/tail_probe.py ) > gpurun_out/t2_probe.txt 2>&1
( TANTE_TAIL_2CTA=0 timeout 200 python tools/tail_probe.py ) > gpurun_out/t2_probe_v1.txt 2>&1
tail -15 gpurun_out/t2_pytest.log; cat gpurun_out/t2_probe.txt gpurun_out/t2_probe_v1.txt
