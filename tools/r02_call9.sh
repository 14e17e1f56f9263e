#!/bin/bash
# GPU job 9 of round 2: parity suite on the current library, per-d sweep timings, the driver's exact N = 1 bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02g.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02g.log
tail -6 gpurun_out/gpu_tests_r02g.log
timeout 300 python tools/dev_time_sweeps.py dims > gpurun_out/dev_time_sweeps_dims_r02g.log 2>&1; echo "sweeps rc=$?"
timeout 300 python tools/dev_time_sweeps.py > gpurun_out/dev_time_sweeps_r02g.log 2>&1; echo "sweeps rc=$?"
T0=$(date +%s)
timeout 870 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver_n1_r02b.out 2> gpurun_out/bench_driver_n1_r02b.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
cat gpurun_out/bench_driver_n1_r02b.err | tail -12
tail -n 1 gpurun_out/bench_driver_n1_r02b.out | cut -c1-400
