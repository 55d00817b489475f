#!/bin/bash
# round-2b experiments: new collection kernel, 2-D cell visibility, tile-order sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_collect_gpu.py tests/test_bake_gpu.py tests/test_visibility_fuzz_gpu.py -x -q > gpurun_out/pytest_try.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_try.log
timeout 300 python tools/sweep_vis.py --config c4 > gpurun_out/sweep_vis_c4.jsonl 2> gpurun_out/sweep_vis_c4.err; echo "sweep_vis rc=$?"; cat gpurun_out/sweep_vis_c4.jsonl
timeout 300 python tools/sweep_collect.py > gpurun_out/sweep_collect_c3.jsonl 2> gpurun_out/sweep_collect_c3.err; echo "sweep_collect rc=$?"; cat gpurun_out/sweep_collect_c3.jsonl
timeout 300 python tools/sweep_collect.py --bands 4 --patches 20000 --samples 2000 --receivers 32 --splits 0 > gpurun_out/sweep_collect_t2000.jsonl 2>> gpurun_out/sweep_collect_c3.err; echo "sweep_collect t2000 rc=$?"; cat gpurun_out/sweep_collect_t2000.jsonl
timeout 400 python tools/sweep_gather.py --config c2 --variants tmem,tmem@natural,tmem@window1184,tmem@window296 > gpurun_out/sweep_gather_c2.log 2>&1; echo "sweep_gather rc=$?"; tail -5 gpurun_out/sweep_gather_c2.log
timeout 300 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_try.json 2> gpurun_out/bench_c3_try.err; echo "bench c3 rc=$?"; cut -c1-300 gpurun_out/bench_c3_try.json
