#!/bin/bash
# final validation of the round: GPU tests, benches, ncu captures of the kernels of round 2b
mkdir -p gpurun_out
M="gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for c in c4 c3; do
  python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/bench_$c.json
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_collect_staged -s 3 -c 1 -f -o gpurun_out/r02_k_collect_staged_c3 python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:"k_vis_p2p_grouped|k_ff_stokes|k_own_in" -c 4 --csv --log-file gpurun_out/r02_bake_kernels_c4.csv python tools/sweep_vis.py --config c4 --reps 1 > gpurun_out/ncu_bake.log 2>&1; echo "ncu bake rc=$?"
python tools/time_d2h.py > gpurun_out/d2h.json 2>&1; cat gpurun_out/d2h.json
