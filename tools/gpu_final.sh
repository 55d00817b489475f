#!/bin/bash
# final validation of the round: GPU tests, benches (both arms), ncu captures of the kernels of round 2b
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for c in c4 c2 c3; do
  python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$? lines=$(wc -l < gpurun_out/bench_$c.json)"; cut -c1-200 gpurun_out/bench_$c.json
done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err; echo "bench reference rc=$? lines=$(wc -l < gpurun_out/bench_ref_c4.json)"; cut -c1-300 gpurun_out/bench_ref_c4.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vis_p2p_grouped -c 1 -f -o gpurun_out/r02_k_vis_p2p_grouped_c4 python tools/sweep_vis.py --config c4 --reps 1 > gpurun_out/ncu_vis.log 2>&1; echo "ncu vis rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ff_stokes -c 1 -f -o gpurun_out/r02_k_ff_stokes_c4 python tools/sweep_vis.py --config c4 --reps 1 > gpurun_out/ncu_ff.log 2>&1; echo "ncu ff rc=$?"
