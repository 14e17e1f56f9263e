#!/bin/bash
# GPU job 24 of round 2: final log of the parity suite and smoke() on HEAD; kin40k-shaped config at the trained operating point
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02q.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02q.log
tail -4 gpurun_out/gpu_tests_r02q.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02q.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02q.log
timeout 200 python bench.py --workload kin40k --theta trained --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kin40k_trained_n1_r02.out 2> gpurun_out/bench_kin40k_trained_n1_r02.err; echo "kin40k trained rc=$?"
tail -n 1 gpurun_out/bench_kin40k_trained_n1_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['roofline']['frac'], j['kv_gpairs_per_s'], j['config']['cg_steps'], j['roofline_other'])"
