#!/bin/bash
# GPU job 6 of round 2: split-phase hand-over of the forward DMMA sweep -- parity, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r02d.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02d.log
tail -6 gpurun_out/gpu_tests_r02d.log
timeout 300 python tools/dev_time_sweeps.py > gpurun_out/dev_time_sweeps_r02d.log 2>&1; echo "sweeps rc=$?"
cat gpurun_out/dev_time_sweeps_r02d.log
timeout 300 python tools/dev_time_sweeps.py dims > gpurun_out/dev_time_sweeps_dims_r02d.log 2>&1; echo "sweeps rc=$?"
cat gpurun_out/dev_time_sweeps_dims_r02d.log
