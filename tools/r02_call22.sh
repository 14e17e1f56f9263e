#!/bin/bash
# GPU job 22 of round 2: bench.py's cumulative-line path (steps > 5 s) and smoke() on HEAD
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02o.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02o.log
timeout 300 python bench.py --n 800000 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n800k_r02.out 2> gpurun_out/bench_n800k_r02.err; echo "bench rc=$?"
grep -c '^{' gpurun_out/bench_n800k_r02.out
python - <<'PY'
import json
l=[json.loads(x) for x in open('gpurun_out/bench_n800k_r02.out') if x.startswith('{')]
for j in l: print(j.get('partial'), j['steps'], round(j['value'],3), j['clocks'], j['roofline']['traffic'])
PY
tail -3 gpurun_out/bench_n800k_r02.err
