#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
( timeout 300 $TR tools/dp_native_check.py ) > gpurun_out/dpq_native_check.log 2>&1; echo "native check exit $?" >> gpurun_out/dpq_native_check.log
timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_r2x_train_n$N.json 2> gpurun_out/dpq_bench.err
tail -2 gpurun_out/dpq_native_check.log; grep -v "^NCCL\|^\*\|OMP_NUM" gpurun_out/bench_r2x_train_n$N.json | head -c 300; echo; tail -3 gpurun_out/dpq_bench.err
