#!/bin/bash
# 8 GPUs: C4 with the whole order fused into one kernel vs gather + mix
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528"
for f in 1 0; do
SPB_FUSED_ORDER=$f timeout 300 $TR bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_c4_8_f$f.json 2> gpurun_out/bench_c4_8_f$f.err; echo "c4 x8 fused=$f rc=$?"; grep "^{" gpurun_out/bench_c4_8_f$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['result']['equals_single_gpu_bitwise'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'])"
done
