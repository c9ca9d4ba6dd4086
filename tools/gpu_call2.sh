#!/bin/bash
# fused block tail bring-up: unit tests first (own process), then the suite, the probe, and short benches
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
( time timeout 600 python -m pytest tests/test_gpu_block_tail.py -q -x ) > gpurun_out/${TAG}_tail_pytest.log 2>&1
RC=$?
echo "tail pytest exit $RC" >> gpurun_out/${TAG}_tail_pytest.log
if [ $RC -ne 0 ]; then export TANTE_FUSE_TAIL=0; echo "FUSE_TAIL disabled for the rest" >> gpurun_out/${TAG}_tail_pytest.log; fi
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_block_tail.py ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
if [ $RC -eq 0 ]; then timeout 300 python tools/tail_probe.py > gpurun_out/${TAG}_tail_probe.txt 2>&1; fi
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/${TAG}_smoke.log 2>&1
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
TANTE_FUSE_TAIL=0 timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras > gpurun_out/${TAG}_bench_nofuse.json 2>> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_tail_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_tail_probe.txt; tail -2 gpurun_out/${TAG}_smoke.log; head -c 600 gpurun_out/${TAG}_bench.json; echo; head -c 600 gpurun_out/${TAG}_bench_nofuse.json; tail -3 gpurun_out/${TAG}_bench.err
