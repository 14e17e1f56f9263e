#!/bin/bash
# GPU job 19 of round 2 (2 GPUs): bench.py's multi-rank path after the parity-check edit, on small workloads
mkdir -p gpurun_out
for w in kin40k 3droad; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --workload $w --steps 3 --warmup 3 > gpurun_out/bench_${w}_n2_r02.out 2> gpurun_out/bench_${w}_n2_r02.err; echo "$w rc=$?"
  tail -n 1 gpurun_out/bench_${w}_n2_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['config']['multi_gpu_parity'])"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --workload snelson1d --steps 2 --warmup 1 | tail -n 1 | cut -c1-200
