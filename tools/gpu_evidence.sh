#!/bin/bash
# Round-2 evidence: final bench lines, ncu launch lists, ncu --set full captures (all outputs under gpurun_out/).
set -u
mkdir -p gpurun_out
TAG=${1:-r2z}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_${TAG}_train_n1.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload rollout > gpurun_out/bench_${TAG}_rollout_n1.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --no-extras --no-eager --no-cpu-baseline --dropout 0.1 > gpurun_out/bench_${TAG}_train_dropout01_n1.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference_train.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --shape viscoelastic --batch 8 --no-extras --no-cpu-baseline > gpurun_out/bench_${TAG}_train_viscoelastic_n1.json 2>> gpurun_out/${TAG}_bench.err
# launch lists (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches_${TAG}_train_b16.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_train.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_bf16_rollout_b64.csv \
    python bench.py --workload rollout --steps 1 --warmup 3 --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_roll.log 2>&1
# full captures in situ (reports stay on the box: only their raw-metric CSV pages come back -- gpurun_out is capped at 64 MiB)
timeout 900 ncu --set full --clock-control none -k regex:'block_tail|gemm_tc|axial_attention_mma|taylor_head' -s 60 -c 8 \
    -o /tmp/${TAG}_full_rollout -f python bench.py --workload rollout --steps 1 --warmup 3 --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_roll.log 2>&1
ncu -i /tmp/${TAG}_full_rollout.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_rollout.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'block_tail|gemm_tc|wgrad_tc|ln_bwd|axial_attention_bwd|mlp_bwd' -s 700 -c 16 \
    -o /tmp/${TAG}_full_train -f python bench.py --steps 1 --warmup 3 --no-extras --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_train.log 2>&1
ncu -i /tmp/${TAG}_full_train.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_train.csv 2>/dev/null
# the forward's block tail (the capture above starts inside the backward)
timeout 600 ncu --set full --clock-control none -k regex:'block_tail' -s 40 -c 2 \
    -o /tmp/${TAG}_full_train_fwd -f python bench.py --steps 1 --warmup 3 --no-extras --no-eager --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_train_fwd.log 2>&1
ncu -i /tmp/${TAG}_full_train_fwd.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_train_fwd.csv 2>/dev/null
du -sh gpurun_out
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log; head -c 300 gpurun_out/bench_${TAG}_train_n1.json; echo; tail -3 gpurun_out/${TAG}_bench.err; ls -la gpurun_out | grep ${TAG}
