#!/bin/bash
# round 2, evidence run after the kernel work: bench line (2 handles), DMMA-pipe counters, launch list, full ncu of the three top kernels, CPU arm
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
M=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i_ncu_stab_dmma.csv python tools/prof_stab.py hubbard_16x16_beta10 148 2 > gpurun_out/r2i_ncu_stab.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -c 140 --csv --log-file gpurun_out/r2i_ncu_taum_dmma.csv python tools/prof_taum.py hubbard_16x16_beta10 148 > gpurun_out/r2i_ncu_taum.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 1 --warmup 1 --handles 1 --no-cpu-baseline > gpurun_out/r2i_launches_bench.json 2> gpurun_out/r2i_launches.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_wrapgr_fast" -c 1 -o gpurun_out/r2i_update python tools/prof_stab.py hubbard_16x16_beta10 148 1 > gpurun_out/r2i_update.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err
