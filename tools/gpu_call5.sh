#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2e}
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python tools/tail_probe.py > gpurun_out/${TAG}_tail_probe.txt 2>&1
timeout 300 python tools/gemm_probe.py > gpurun_out/${TAG}_gemm_probe.txt 2>&1
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras --dropout 0.1 > gpurun_out/${TAG}_bench_p01.json 2>> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_tail_probe.txt; tail -12 gpurun_out/${TAG}_gemm_probe.txt; head -c 400 gpurun_out/${TAG}_bench.json; echo; head -c 300 gpurun_out/${TAG}_bench_p01.json; echo; tail -3 gpurun_out/${TAG}_bench.err
