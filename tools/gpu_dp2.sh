#!/bin/bash
# 2-GPU: native all-reduce check, default bench line (training + rollout records), viscoelastic weak / strong
mkdir -p gpurun_out
TAG=${1:-dp2}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( timeout 300 $TR tools/dp_native_check.py ) > gpurun_out/${TAG}_native_check.log 2>&1
echo "native check exit $?" >> gpurun_out/${TAG}_native_check.log
timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_${TAG}_train_n$N.json 2> gpurun_out/${TAG}_bench.err
timeout 600 $TR bench.py --gpus $N --shape viscoelastic --batch 8 --no-extras > gpurun_out/bench_${TAG}_train_viscoelastic_weak_n$N.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 $TR bench.py --gpus $N --shape viscoelastic --global-batch 64 --no-extras > gpurun_out/bench_${TAG}_train_viscoelastic_strong_n$N.json 2>> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_native_check.log
for f in gpurun_out/bench_${TAG}_*_n$N.json; do echo $f; head -c 400 $f; echo; done
tail -5 gpurun_out/${TAG}_bench.err
