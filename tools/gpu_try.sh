#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests/test_collect_gpu.py tests/test_class_gpu.py -x -q > gpurun_out/pytest_try.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_try.log
timeout 300 python tools/sweep_collect.py --splits 0 > gpurun_out/sweep_collect_final.jsonl 2>&1; cat gpurun_out/sweep_collect_final.jsonl
timeout 300 python bench.py --config c1 --steps 3 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 rc=$?"; cut -c1-300 gpurun_out/bench_c1.json
