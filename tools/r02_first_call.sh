#!/bin/bash
# First GPU job of a round (run as:  gpurun --timeout 1500 -- 'bash tools/r02_first_call.sh r02'):
#   1. GPU parity suite, 2. default bench line (n = 2M, one B200; ~6 min), 3. launch list of the n = 400k bench
#   command (the n = 2M command does not fit the pool's ncu wrapper), 4. one ncu --set full capture of the
#   dominant kernel (DMMA K*v sweep, d = 11), each ncu pass only after its own command exited 0 without ncu.
# Outputs go to gpurun_out/ with the round tag; copy what is to be judged into profiles/.
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_${TAG}.log
tail -3 gpurun_out/gpu_tests_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_default_${TAG}.json 2> gpurun_out/bench_default_${TAG}.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_default_${TAG}.json
timeout 200 python bench.py --n 400000 --no-cpu-baseline > gpurun_out/bench_n400k_${TAG}.json 2> gpurun_out/bench_n400k_${TAG}.err \
  && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
       --log-file gpurun_out/launches_bench_houseelectric_n400k_${TAG}.csv python bench.py --n 400000 --no-cpu-baseline \
       > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 120 python tools/prof_kmv.py matern32 200000 11 2 > gpurun_out/prof_plain_${TAG}.log 2>&1 \
  && timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmma_sweep_kernel -c 1 \
       -o gpurun_out/dsweep_d11_${TAG} -f python tools/prof_kmv.py matern32 200000 11 2 > gpurun_out/ncu_dsweep_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
