#!/bin/bash
# GPU job 10 of round 2: the TMA-staged GEMM -- parity (dense, config sizes, both staging paths), then dense timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sizes.py -m gpu -q -x > gpurun_out/gpu_tests_r02h.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02h.log
tail -8 gpurun_out/gpu_tests_r02h.log
timeout 400 python tools/dev_dense_time.py > gpurun_out/dev_dense_time_r02h.log 2>&1; echo "dense rc=$?"
cat gpurun_out/dev_dense_time_r02h.log
