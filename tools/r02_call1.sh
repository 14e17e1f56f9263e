#!/bin/bash
# GPU job 1 of round 2: parity suite, gradient-spread diagnostic, the driver's exact bench command under its 870 s limit.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r02a.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02a.log
tail -3 gpurun_out/gpu_tests_r02a.log
timeout 200 python tools/grad_spread_gpu.py > gpurun_out/grad_spread_gpu_r02.md 2> gpurun_out/grad_spread_gpu_r02.err; echo "spread rc=$?"
tail -4 gpurun_out/grad_spread_gpu_r02.md
/usr/bin/time -v timeout 870 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver_n1_r02.out 2> gpurun_out/bench_driver_n1_r02.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_driver_n1_r02.err
tail -n 1 gpurun_out/bench_driver_n1_r02.out | cut -c1-1500
