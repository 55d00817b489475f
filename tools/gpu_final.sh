#!/bin/bash
# final validation of the round: GPU tests, benches (both arms), launch list, ncu captures
mkdir -p gpurun_out
M="gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for c in c4 c2 c3; do
  python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"; cut -c1-200 gpurun_out/bench_$c.json
done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err; echo "bench reference rc=$?"; cut -c1-400 gpurun_out/bench_ref_c4.json
SPB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gather_tmem -s 2 -c 1 -f -o gpurun_out/r02_k_gather_tmem_c4 python tools/profile_gather.py --config c4 --gather tmem > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_vis_p2p_grouped|k_ff_stokes|k_ff_nusselt|k_pair_geometry" -c 4 --csv --log-file gpurun_out/r02_bake_kernels_c4.csv python tools/profile_gather.py --config c4 --gather tmem --orders 1 > gpurun_out/ncu_bake.log 2>&1; echo "ncu bake rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_mix" -s 1 -c 1 --csv --log-file gpurun_out/r02_k_mix_c2.csv python tools/profile_gather.py --config c2 --gather tmem --orders 2 > gpurun_out/ncu_mix.log 2>&1; echo "ncu mix rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_collect_partial|k_source_energy|k_receiver_factors|k_init_scatter" -c 4 --csv --log-file gpurun_out/r02_c3_kernels.csv python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
