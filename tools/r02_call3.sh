#!/bin/bash
# GPU job 3 of round 2: parity after the fixed-order reductions, gradient spread, sweep timings, n = 400k bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r02b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02b.log
tail -15 gpurun_out/gpu_tests_r02b.log
timeout 200 python tools/grad_spread_gpu.py > gpurun_out/grad_spread_gpu_r02b.md 2> gpurun_out/grad_spread_gpu_r02b.err; echo "spread rc=$?"
tail -4 gpurun_out/grad_spread_gpu_r02b.md
timeout 300 python tools/dev_time_sweeps.py > gpurun_out/dev_time_sweeps_r02b.log 2>&1; echo "sweeps rc=$?"
cat gpurun_out/dev_time_sweeps_r02b.log
timeout 300 python bench.py --n 400000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n400k_r02b.out 2> gpurun_out/bench_n400k_r02b.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n400k_r02b.err; tail -n 1 gpurun_out/bench_n400k_r02b.out | cut -c1-600
