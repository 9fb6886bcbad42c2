#!/bin/bash
# round 2, final evidence of the shipped build: full GPU test suite, smoke, bench (1 GPU), launch list, CPU arm
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 1 --warmup 1 --handles 1 --no-cpu-baseline > gpurun_out/r2z_launches_bench.json 2> gpurun_out/r2z_launches.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err
tail -3 gpurun_out/r2z_pytest_gpu.log; tail -2 gpurun_out/r2z_smoke.log; head -c 300 gpurun_out/r2z_bench.json; echo; head -c 200 gpurun_out/r2z_ref.json
