#!/bin/bash
# GPU job 4 of round 2: full parity suite (fixed-order reductions, multi-RHS, Nystrom-size tests), gradient spread
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02c.log
tail -25 gpurun_out/gpu_tests_r02c.log
timeout 200 python tools/grad_spread_gpu.py > gpurun_out/grad_spread_gpu_r02c.md 2> gpurun_out/grad_spread_gpu_r02c.err; echo "spread rc=$?"
tail -4 gpurun_out/grad_spread_gpu_r02c.md
