#!/bin/bash
# GPU job 21 of round 2: super-row size of the DMMA sweeps' item order vs DRAM traffic and time at n = 2M (metrics-only ncu)
mkdir -p gpurun_out/ncu
for q in 64 128 256; do
  export CGLB_SUPERROW=$q
  python tools/prof_kmv_fwd.py matern32 2000000 11 > gpurun_out/ncu/dsweep_n2M_q${q}_plain.log 2>&1 && timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dmma_sweep_kernel -c 1 --csv --log-file gpurun_out/ncu/dsweep_d11_n2M_q${q}_dram.csv python tools/prof_kmv_fwd.py matern32 2000000 11 > gpurun_out/ncu/dsweep_n2M_q${q}_ncu.log 2>&1; echo "q=$q ncu rc=$?"
  cat gpurun_out/ncu/dsweep_n2M_q${q}_plain.log; grep -o '"dram__bytes_[a-z]*.sum","byte","[0-9]*"\|"lts__t_sector_hit_rate.pct","%","[0-9.]*"' gpurun_out/ncu/dsweep_d11_n2M_q${q}_dram.csv
done
