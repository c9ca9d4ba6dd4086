#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2f}
( time timeout 1500 python -m pytest tests -m gpu -q -x -k "next_scope or refused" ) > gpurun_out/${TAG}_next_pytest.log 2>&1
echo "next pytest exit $?" >> gpurun_out/${TAG}_next_pytest.log
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_next_scope_forward_and_rollout_fp32 --deselect tests/test_gpu_parity.py::test_next_scope_forward_bf16 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_next_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log; head -c 300 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
