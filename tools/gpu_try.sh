#!/bin/bash
# round-2b experiment: 512-thread shape of the staged collection
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_collect_gpu.py -x -q > gpurun_out/pytest_try.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_try.log
timeout 300 python tools/sweep_collect.py --variants staged:2:3,staged:4:2,staged:4:3,staged:4:4,staged:1:3 --splits 0,16,32 > gpurun_out/sweep_collect_c3b.jsonl 2> gpurun_out/sweep_collect_c3b.err; echo "sweep_collect rc=$?"; cat gpurun_out/sweep_collect_c3b.jsonl
timeout 300 python tools/sweep_collect.py --bands 4 --patches 20000 --samples 2000 --receivers 32 --splits 0 --variants staged:2:3,staged:4:2,staged:4:3 > gpurun_out/sweep_collect_t2000b.jsonl 2>> gpurun_out/sweep_collect_c3b.err; echo "sweep_collect t2000 rc=$?"; cat gpurun_out/sweep_collect_t2000b.jsonl
