#!/bin/bash
# GPU job 13 of round 2 (8 GPUs): the driver's launch line for the headline config with a short budget, and the 3droad-shaped
# config (BASELINE.json configs[2]: "1/2/4/8 B200")
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # name, extra args...
  local name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 "$@" > gpurun_out/bench_${name}_n8_r02.out 2> gpurun_out/bench_${name}_n8_r02.err; echo "$name rc=$?"
  grep "\[bench\]" gpurun_out/bench_${name}_n8_r02.err | tail -4
  tail -n 1 gpurun_out/bench_${name}_n8_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['warmup'], j['config'].get('multi_gpu_parity',{}).get('ok'), j['roofline']['frac'], j['config']['cg_steps'])"
}
run houseelectric --steps 20 --warmup 5 --max-seconds 230
run 3droad --workload 3droad --steps 6 --warmup 3
run kin40k --workload kin40k --steps 6 --warmup 3
