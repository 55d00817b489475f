#!/bin/bash
# 8 GPUs: config 5 (100k patches, band-wise schedule, sharded bake), C4 scaling at 8 and 4
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 900 $TR8 bench.py --config c5 --gpus 8 --steps 2 --warmup 3 --verbose > gpurun_out/bench_c5_8.json 2> gpurun_out/bench_c5_8.err; echo "c5 x8 rc=$?"; grep "bench\]\|Error\|error" gpurun_out/bench_c5_8.err | tail -5; grep "^{" gpurun_out/bench_c5_8.json | cut -c1-600
timeout 300 $TR8 bench.py --gpus 8 > gpurun_out/bench_c4_8.json 2> gpurun_out/bench_c4_8.err; echo "c4 x8 rc=$?"; grep "^{" gpurun_out/bench_c4_8.json | cut -c1-300
timeout 300 $TR4 bench.py --gpus 4 > gpurun_out/bench_c4_4.json 2> gpurun_out/bench_c4_4.err; echo "c4 x4 rc=$?"; grep "^{" gpurun_out/bench_c4_4.json | cut -c1-300
nvidia-smi --query-gpu=memory.used --format=csv | head -3
