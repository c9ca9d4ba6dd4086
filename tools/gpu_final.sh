#!/bin/bash
# Final evidence of round 2 (tag r2f): smoke(), the default bench line, the rollout line, the reference arm, launch lists of both workloads.
mkdir -p gpurun_out
TAG=${1:-r2f}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_train_n1.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload rollout > gpurun_out/bench_${TAG}_rollout_n1.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_train.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}_train_b16.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_train.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_bf16_rollout_b64.csv \
    python bench.py --workload rollout --steps 1 --warmup 3 --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_roll.log 2>&1
head -c 400 gpurun_out/bench_${TAG}_train_n1.json; echo; head -c 300 gpurun_out/bench_${TAG}_rollout_n1.json; echo; head -c 300 gpurun_out/bench_${TAG}_reference_train.json; echo; tail -2 gpurun_out/${TAG}_bench.err
