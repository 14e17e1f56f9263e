#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "m256" > gpurun_out/gpu_tests_r02p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_r02p.log
timeout 100 python tools/grad_spread_gpu.py 2>/dev/null | grep "m256\|worst"
