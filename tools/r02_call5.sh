#!/bin/bash
# GPU job 5 of round 2 (2 GPUs): row-sharded parity test + the driver's launch line at N = 2 with a short budget
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/gpu_multi_r02.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_multi_r02.log
tail -5 gpurun_out/gpu_multi_r02.log
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --max-seconds 330 > gpurun_out/bench_driver_n2_r02.out 2> gpurun_out/bench_driver_n2_r02.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
grep "\[bench\]" gpurun_out/bench_driver_n2_r02.err | tail -12
tail -n 1 gpurun_out/bench_driver_n2_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['warmup'], j['config']['multi_gpu_parity'], j['roofline']['frac'], j['gpu_launches'])"
