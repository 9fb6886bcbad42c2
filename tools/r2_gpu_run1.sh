#!/bin/bash
# round 2, GPU call 1: headline-size parity tests, bench line, DMMA-pipe counters of the dense kernels
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 1500 python -m pytest tests/test_gpu_headline_sizes.py -q -m gpu --timeout=900 > gpurun_out/r2a_pytest_headline.log 2>&1
echo "pytest headline rc=$?" >> gpurun_out/r2a_pytest_headline.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
M=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2a_ncu_stab_dmma.csv python tools/prof_stab.py hubbard_16x16_beta10 148 2 > gpurun_out/r2a_ncu_stab.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -c 140 --csv --log-file gpurun_out/r2a_ncu_taum_dmma.csv python tools/prof_taum.py hubbard_16x16_beta10 148 > gpurun_out/r2a_ncu_taum.log 2>&1
