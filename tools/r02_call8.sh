#!/bin/bash
# GPU job 8 of round 2: remaining ncu evidence at SHORT shapes (each --set full pass replays the kernel ~40 times) and the
# launch list of a kin40k-shaped bench command.  Reports are exported to CSV on the box and deleted (64 MiB pull limit).
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
NCU="ncu --set full --clock-control none"
run() {  # name, regex, count, cmd...
  local name=$1 regex=$2 cnt=$3; shift 3
  "$@" > $O/${name}_plain.log 2>&1 && timeout 240 $NCU -k regex:$regex -c $cnt -o $O/$name -f "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"; tail -1 $O/${name}_ncu.log
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null; rm -f $O/$name.ncu-rep
}
run kmv_rbf_d8_n40k "kmv_sweep_kernel|dmma_bwd_kernel|kmv_bwd_kernel" 2 python tools/prof_kmv.py rbf 40000 8 1
run wide_d90_n100k "wide_sweep_kernel|wide_bwd_kernel" 2 python tools/prof_kmv.py matern32 100000 90 1
run dense_m2048_n100k "gemv_rows_kernel|gemv_cols_finish_kernel|gemm_kernel" 14 python tools/prof_dense.py 2048 100000
python bench.py --workload kin40k --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_kin40k_plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench_kin40k_r02.csv python bench.py --workload kin40k --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_kin40k_ncu.log 2>&1; echo "launch list rc=$?"
ls -la $O; du -sh gpurun_out
