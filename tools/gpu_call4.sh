#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2d}
( time timeout 900 python -m pytest tests/test_gpu_dropout.py -q ) > gpurun_out/${TAG}_dropout_pytest.log 2>&1
echo "dropout pytest exit $?" >> gpurun_out/${TAG}_dropout_pytest.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dropout.py ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras > gpurun_out/${TAG}_bench_p0.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras --dropout 0.1 > gpurun_out/${TAG}_bench_p01.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload rollout --no-cpu-baseline > gpurun_out/${TAG}_bench_roll.json 2>> gpurun_out/${TAG}_bench.err
tail -15 gpurun_out/${TAG}_dropout_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log; head -c 300 gpurun_out/${TAG}_bench_p0.json; echo; head -c 300 gpurun_out/${TAG}_bench_p01.json; echo; tail -3 gpurun_out/${TAG}_bench.err
