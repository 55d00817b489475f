#!/bin/bash
# fused order: parity (1 GPU), sharded parity + timing at 2 GPUs
mkdir -p gpurun_out
python -m pytest tests/test_exchange_gpu.py tests/test_window_gather_gpu.py tests/test_class_gpu.py tests/test_fullsize_c4_gpu.py -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench_c4_fused.json 2> gpurun_out/bench_c4_fused.err; echo "bench c4 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_fused.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['gpu_launches'], d['result'])"
SPB_FUSED_ORDER=0 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_unfused.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_unfused.json')); print('unfused', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['result'])"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
CHECK_CONFIG=c4 timeout 600 $TR tools/check_sharded.py > gpurun_out/check_sharded_c4_2.log 2>&1; echo "check_sharded c4 rc=$?"; grep "equal\|PARITY\|Error\|error" gpurun_out/check_sharded_c4_2.log | tail -20
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_c4_2.json 2> gpurun_out/bench_c4_2.err; echo "c4 x2 rc=$?"; grep "^{" gpurun_out/bench_c4_2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['result'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'])"
