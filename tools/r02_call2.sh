#!/bin/bash
# GPU job 2 of round 2: the driver's exact bench command under its 870 s limit (N = 1).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 870 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver_n1_r02.out 2> gpurun_out/bench_driver_n1_r02.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -c 800 gpurun_out/bench_driver_n1_r02.err
grep -c '^{' gpurun_out/bench_driver_n1_r02.out
tail -n 1 gpurun_out/bench_driver_n1_r02.out | cut -c1-1800
