#!/bin/bash
# round-2b experiments: collection kernel with asynchronous receiver factors, 16-lane Stokes kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_collect_gpu.py tests/test_bake_gpu.py tests/test_fullsize_c4_gpu.py tests/test_class_gpu.py -x -q > gpurun_out/pytest_try.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_try.log
timeout 300 python tools/sweep_collect.py --splits 0,16 > gpurun_out/sweep_collect_c3.jsonl 2> gpurun_out/sweep_collect_c3.err; echo "sweep_collect rc=$?"; cat gpurun_out/sweep_collect_c3.jsonl
timeout 300 python tools/sweep_collect.py --bands 4 --patches 20000 --samples 2000 --receivers 32 --splits 0 > gpurun_out/sweep_collect_t2000.jsonl 2>> gpurun_out/sweep_collect_c3.err; echo "sweep_collect t2000 rc=$?"; cat gpurun_out/sweep_collect_t2000.jsonl
timeout 400 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_try.json 2> gpurun_out/bench_c4_try.err; echo "bench c4 rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_c4_try.json'))
print(d['ms_per_step'], d['bake'], d['pipeline'])
P
