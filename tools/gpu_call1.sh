#!/bin/bash
# round-2 GPU call: tests, bench C4/C2, launch list, ncu full capture of the gather, microbench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log
python bench.py --verbose > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
cat gpurun_out/bench_c4.json | cut -c1-1500
python bench.py --config c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
cut -c1-600 gpurun_out/bench_c2.json
tools/microbench/ub > gpurun_out/ub.jsonl 2>&1; echo "ub rc=$?"
head -5 gpurun_out/ub.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gather_tmem -s 2 -c 1 -f -o gpurun_out/r02_k_gather_tmem_c4 python tools/profile_gather.py --config c4 --gather tmem > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full.log
