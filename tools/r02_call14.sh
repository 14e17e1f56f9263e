#!/bin/bash
# GPU job 14 of round 2: final validation of HEAD on one B200 -- parity suite, smoke(), the small configs' bench lines, then
# the default bench command (no flags)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02j.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02j.log
tail -4 gpurun_out/gpu_tests_r02j.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02j.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_r02j.log
for w in snelson1d kin40k 3droad song; do
  timeout 300 python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/bench_${w}_n1_r02.out 2> gpurun_out/bench_${w}_n1_r02.err; echo "$w rc=$?"
  tail -n 1 gpurun_out/bench_${w}_n1_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['roofline']['frac'], j['kv_gpairs_per_s'], j['config']['cg_steps'], j['cpu_baseline']['value'])"
done
timeout 300 python bench.py --workload 3droad --theta trained --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_3droad_trained_n1_r02.out 2> gpurun_out/bench_3droad_trained_n1_r02.err; echo "3droad trained rc=$?"
tail -n 1 gpurun_out/bench_3droad_trained_n1_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['roofline']['frac'], j['kv_gpairs_per_s'], j['config']['cg_steps'])"
T0=$(date +%s)
timeout 870 python bench.py > gpurun_out/bench_default_r02.out 2> gpurun_out/bench_default_r02.err; echo "default bench rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -5 gpurun_out/bench_default_r02.err
tail -n 1 gpurun_out/bench_default_r02.out | cut -c1-300
