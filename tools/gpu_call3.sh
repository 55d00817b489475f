#!/bin/bash
# 2-GPU checks: sharded parity in every comm mode, C5-small band-wise schedule, C4 at 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/check_sharded.py > gpurun_out/check_sharded_2.log 2>&1; echo "check_sharded rc=$?"; grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/check_sharded_2.log | tail -25
timeout 600 python bench.py --config c5s --no-cpu-baseline > gpurun_out/bench_c5s_1.json 2> gpurun_out/bench_c5s_1.err; echo "c5s x1 rc=$?"; cut -c1-400 gpurun_out/bench_c5s_1.json; tail -3 gpurun_out/bench_c5s_1.err
timeout 600 $TR bench.py --config c5s --gpus 2 > gpurun_out/bench_c5s_2.json 2> gpurun_out/bench_c5s_2.err; echo "c5s x2 rc=$?"; cut -c1-400 gpurun_out/bench_c5s_2.json; tail -3 gpurun_out/bench_c5s_2.err
timeout 600 $TR bench.py --gpus 2 > gpurun_out/bench_c4_2.json 2> gpurun_out/bench_c4_2.err; echo "c4 x2 rc=$?"; cut -c1-300 gpurun_out/bench_c4_2.json; tail -3 gpurun_out/bench_c4_2.err
