#!/bin/bash
# GPU job 16 of round 2: HEAD after the GEMM default went back to the cp.async ring -- parity suite, smoke(), small configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02l.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02l.log
tail -4 gpurun_out/gpu_tests_r02l.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02l.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02l.log
for w in snelson1d kin40k 3droad song; do
  timeout 300 python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/bench_${w}_n1_r02.out 2> gpurun_out/bench_${w}_n1_r02.err; echo "$w rc=$?"
  tail -n 1 gpurun_out/bench_${w}_n1_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['roofline']['frac'], j['kv_gpairs_per_s'], j['config']['cg_steps'], j['cpu_baseline']['value'], j['roofline_other']['dense_trsm_syrk_gemm'])"
done
timeout 300 python tools/dev_dense_time.py > gpurun_out/dev_dense_time_r02l.log 2>&1; cat gpurun_out/dev_dense_time_r02l.log
