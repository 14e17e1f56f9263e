#!/bin/bash
# GPU job 18 of round 2 (4 GPUs): the 3droad-shaped config (BASELINE.json configs[2]: "1/2/4/8 B200") at both operating points
mkdir -p gpurun_out
run() { # name, extra args...
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 "$@" > gpurun_out/bench_${name}_n4_r02.out 2> gpurun_out/bench_${name}_n4_r02.err; echo "$name rc=$?"
  tail -n 1 gpurun_out/bench_${name}_n4_r02.out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['steps'], j['warmup'], j['config'].get('multi_gpu_parity',{}).get('ok'), j['roofline']['frac'], j['config']['cg_steps'])"
}
run 3droad_init --workload 3droad --steps 6 --warmup 3
run 3droad_trained --workload 3droad --theta trained --steps 3 --warmup 2
