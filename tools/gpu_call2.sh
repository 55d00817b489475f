#!/bin/bash
# tests, bench C4/C2/C3, launch list (timed steps only), ncu full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log
for c in c4 c2 c3; do
  python bench.py --config $c --verbose > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"
  cut -c1-300 gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err
done
SPB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
for c in c4 c2; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gather_tmem -s 2 -c 1 -f -o gpurun_out/r02_k_gather_tmem_$c python tools/profile_gather.py --config $c --gather tmem > gpurun_out/ncu_full_$c.log 2>&1; echo "ncu full $c rc=$?"
done
