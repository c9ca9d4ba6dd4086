#!/bin/bash
# Refresh of the headline lines after the last code changes (tag r2x): default bench line (traffic filled from profiles/traffic_r2w.json),
# rollout line, rollout launch list, the forward block tail of the training step under ncu --set full.
mkdir -p gpurun_out
TAG=${1:-r2x}
timeout 900 python bench.py > gpurun_out/bench_${TAG}_train_n1.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload rollout > gpurun_out/bench_${TAG}_rollout_n1.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_bf16_rollout_b64.csv \
    python bench.py --workload rollout --steps 1 --warmup 3 --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_roll.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'block_tail' -s 40 -c 2 \
    -o /tmp/${TAG}_full_train_fwd -f python bench.py --steps 1 --warmup 3 --no-extras --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_train_fwd.log 2>&1
ncu -i /tmp/${TAG}_full_train_fwd.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_train_fwd.csv 2>/dev/null
head -c 300 gpurun_out/bench_${TAG}_train_n1.json; echo; tail -2 gpurun_out/${TAG}_bench.err
