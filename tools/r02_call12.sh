#!/bin/bash
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
python tools/prof_gemm.py 2048 100000 > $O/gemm_big_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o $O/gemm_big -f python tools/prof_gemm.py 2048 100000 > $O/gemm_big_ncu.log 2>&1; echo "rc=$?"
ncu -i $O/gemm_big.ncu-rep --page raw --csv > $O/gemm_m2048_n100k_raw.csv 2>/dev/null
ncu -i $O/gemm_big.ncu-rep --page source --csv > $O/gemm_m2048_n100k_source.csv 2>/dev/null
ls -la $O; rm -f $O/gemm_big.ncu-rep
