#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02k.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02k.log
for mode in 1 2; do
  for w in snelson1d kin40k; do
    CGLB_GEMM_STAGING=$mode timeout 200 python bench.py --workload $w --steps 9 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('staging $mode', '$w', j['value'], j['roofline_other']['dense_trsm_syrk_gemm'])"
  done
done
CGLB_GEMM_STAGING=1 timeout 300 python tools/dev_dense_time.py 2>&1 | grep -v cublas | grep "1024\|54250"
