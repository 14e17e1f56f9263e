#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02m.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02m.log
tail -4 gpurun_out/gpu_tests_r02m.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02m.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --workload song --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('song', j['value'], j['kv_gpairs_per_s'])"
