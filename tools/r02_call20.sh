#!/bin/bash
# GPU job 20 of round 2: L2-blocked item order of the DMMA sweeps -- parity, timing, DRAM traffic (metrics-only ncu pass)
mkdir -p gpurun_out/ncu
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02n.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02n.log
tail -4 gpurun_out/gpu_tests_r02n.log
timeout 200 python tools/dev_time_sweeps.py 2>&1 | head -4
python tools/prof_kmv_fwd.py matern32 2000000 11 > gpurun_out/ncu/dsweep_n2M_l2_plain.log 2>&1 && timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dmma_sweep_kernel -c 1 --csv --log-file gpurun_out/ncu/dsweep_d11_n2M_l2blocked_dram.csv python tools/prof_kmv_fwd.py matern32 2000000 11 > gpurun_out/ncu/dsweep_n2M_l2_ncu.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/ncu/dsweep_n2M_l2_plain.log; tail -8 gpurun_out/ncu/dsweep_d11_n2M_l2blocked_dram.csv
