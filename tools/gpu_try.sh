#!/bin/bash
# round-2b experiments: visibility hints (own-polygon memo, own walls first), Stokes without sqrt, c3 bench
mkdir -p gpurun_out
timeout 300 python tools/sweep_vis.py --config c4 > gpurun_out/sweep_vis_c4.jsonl 2> gpurun_out/sweep_vis_c4.err; echo "sweep_vis rc=$?"; cat gpurun_out/sweep_vis_c4.jsonl; tail -2 gpurun_out/sweep_vis_c4.err
timeout 900 python -m pytest tests/test_bake_gpu.py tests/test_visibility_fuzz_gpu.py tests/test_fullsize_c4_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/pytest_try.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_try.log
timeout 300 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_try.json 2> gpurun_out/bench_c3_try.err; echo "bench c3 rc=$?"; cut -c1-400 gpurun_out/bench_c3_try.json
