#!/bin/bash
# First GPU job: FP64 denominators (DFMA loop, DMMA loop, cuBLAS DGEMM) + clocks.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/peaks_clocks.csv &
SMI=$!
./tools/fp64_peak > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err
cat gpurun_out/fp64_peak.json
python - <<'PY' > gpurun_out/dgemm_peak.json 2>&1
import torch, json, time, os
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 0
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): c = a @ b
e1.record(); torch.cuda.synchronize()
sus = 20 * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps({"cublas_dgemm_tflops_burst": best, "cublas_dgemm_tflops_sustained": sus, "cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}))
PY
cat gpurun_out/dgemm_peak.json
kill $SMI
nproc; free -g | head -2; lscpu | head -20 > gpurun_out/lscpu.txt
