#!/bin/bash
# GPU job 25 of round 2: ncu --set full of the preconditioner GEMV pair (K3/K4) at M = 2048, n = 400k
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
python tools/prof_dense.py 2048 400000 > $O/gemv_plain.log 2>&1 && timeout 200 ncu --set full --clock-control none -k regex:"gemv_rows_kernel|gemv_cols_finish_kernel" -s 2 -c 2 -o $O/gemv_pair -f python tools/prof_dense.py 2048 400000 > $O/gemv_ncu.log 2>&1; echo "rc=$?"
ncu -i $O/gemv_pair.ncu-rep --page raw --csv > $O/gemv_pair_m2048_n400k_raw.csv 2>/dev/null; rm -f $O/gemv_pair.ncu-rep
ls -la $O | tail -5
