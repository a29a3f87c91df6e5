#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench line, ncu launch list.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS} 2>&1 | tail -3 | tee gpurun_out/bench.log
if [ -n "$DO_NCU" ]; then
  echo "== ncu launches"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
fi
