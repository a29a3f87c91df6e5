#!/bin/bash
# One multi-GPU box (G = number of visible GPUs): config-4 sweep subset sharded over G ranks, config-5 grasp driver, bench.py at G.
mkdir -p gpurun_out
G=${G:-$(nvidia-smi -L | wc -l)}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541"
echo "== sweep (config 4 subset, N = 16M over $G GPUs)"
timeout 600 $TR scripts/bench_sweep_full.py --iters 2 --shapes ${SHAPES:-512:64,4096:64,16384:64,4096:128,16384:128,16384:256,4096:512} 2>&1 | grep "^{" | cut -c1-330
echo "== grasp (config 5: 10k objects x 100 grasps over $G GPUs)"
timeout 600 $TR scripts/bench_grasp_dist.py 2>&1 | grep "^{" | cut -c1-500
echo "== bench.py --gpus $G"
timeout 600 $TR bench.py --gpus $G --steps 20 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench_n$G.log | cut -c1-300
