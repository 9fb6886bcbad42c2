#!/bin/bash
# round 2: bench lines of every BASELINE configuration (1 GPU) with the current build
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
timeout 600 python bench.py --workload hubbard_8x8_beta10 --steps 5 --warmup 3 > gpurun_out/r2n_bench_8x8.json 2> gpurun_out/r2n_bench_8x8.err
timeout 600 python bench.py --workload hubbard_4x4_beta5 --steps 10 --warmup 3 > gpurun_out/r2n_bench_4x4.json 2> gpurun_out/r2n_bench_4x4.err
timeout 900 python bench.py --workload kondo_12x12_beta20 --steps 1 --warmup 3 --ltau 0 > gpurun_out/r2n_bench_kondo.json 2> gpurun_out/r2n_bench_kondo.err
timeout 900 python bench.py --workload z2_matter_12x12 --steps 1 --warmup 3 --ltau 0 --handles 1 > gpurun_out/r2n_bench_z2.json 2> gpurun_out/r2n_bench_z2.err
for f in gpurun_out/r2n_bench*.json; do echo $f; head -c 400 $f; echo; done
