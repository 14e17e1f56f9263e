#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or dense" > gpurun_out/gpu_tests_r02i.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_r02i.log
timeout 400 python tools/dev_dense_time.py > gpurun_out/dev_dense_time_r02i.log 2>&1; echo "dense rc=$?"
grep -v cublas gpurun_out/dev_dense_time_r02i.log
