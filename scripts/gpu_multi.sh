#!/bin/bash
# N-GPU leg: dist unit test on GPUs + bench at N ranks (torchrun), outputs -> gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/smi_multi.txt
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 2>&1 | tail -4 | tee gpurun_out/bench_n$N.log
echo "== dist parity N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    scripts/dist_parity.py 2>&1 | tail -6 | tee gpurun_out/dist_parity_n$N.log
