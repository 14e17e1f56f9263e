#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "predict" > gpurun_out/gpu_tests_r02r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_r02r.log
