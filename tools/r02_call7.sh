#!/bin/bash
# GPU job 7 of round 2: ncu evidence (one ncu family per call): --set full of the dominant kernels at the configs' own
# shapes + the launch list of a bench command.  Every command first runs plain (&& directly before ncu).
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
NCU="ncu --set full --clock-control none"
run() {  # name, regex, count, cmd...
  local name=$1 regex=$2 cnt=$3; shift 3
  "$@" > $O/${name}_plain.log 2>&1 && $NCU -k regex:$regex -c $cnt -o $O/$name -f "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"; tail -2 $O/${name}_ncu.log
}
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests_r02e.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_tests_r02e.log
timeout 200 python tools/dev_time_sweeps.py > gpurun_out/dev_time_sweeps_r02e.log 2>&1; head -4 gpurun_out/dev_time_sweeps_r02e.log
# headline kernel at the headline shape (n = 2M, d = 11): one launch, ~40 replays of 4.2 s
python tools/prof_kmv.py matern32 2000000 11 1 > $O/dsweep_n2M_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmma_sweep_kernel -c 1 -o $O/dsweep_d11_n2M -f python tools/prof_kmv.py matern32 2000000 11 1 > $O/dsweep_d11_n2M_ncu.log 2>&1; echo "dsweep n2M rc=$?"; tail -2 $O/dsweep_d11_n2M_ncu.log
run dbwd_d11_n1M dmma_bwd_kernel 1 python tools/prof_kmv.py matern32 1000000 11 1
run kmv_d3_n434k "kmv_sweep_kernel|kmv_bwd_kernel" 2 python tools/prof_kmv.py matern32 434000 3 1
run kmv_rbf_d8_n40k "kmv_sweep_kernel|kmv_bwd_kernel|dmma_bwd_kernel" 2 python tools/prof_kmv.py rbf 40000 8 1
run wide_d90_n515k "wide_sweep_kernel|wide_bwd_kernel" 2 python tools/prof_kmv.py matern32 515000 90 1
run dense_m2048_n400k "gemv_rows_kernel|gemv_cols_finish_kernel|gemm_kernel" 12 python tools/prof_dense.py 2048 400000
# launch list of a bench command (n = 400k: the n = 2M command does not fit the plain-run limit of the ncu wrapper)
python bench.py --n-rows 400000 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_n400k_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench_houseelectric_n400k_r02.csv python bench.py --n-rows 400000 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_n400k_ncu.log 2>&1; echo "launch list rc=$?"
ls -la $O | head -40; du -sh $O
