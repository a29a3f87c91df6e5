#!/bin/bash
# streamed-codebook kernel: default and the DVQ_TC_ST / DVQ_TC_CE variants: canary (streamed shape), tc vs FP32 kernel on many shapes, VQ GPU tests, sweep subset, bench
mkdir -p gpurun_out
N=262144 K=4096 D=128 ITERS=2 timeout 90 python scripts/tc_profile.py 2>&1 | tail -2 | tee gpurun_out/canary.log; grep -q "^loss" gpurun_out/canary.log || { echo CANARY FAILED; exit 1; }
timeout 300 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo rc=$?
cut -c1-330 gpurun_out/tc_debug.log | grep -o "^([0-9, a-z_']*)\|tc_ms.: [0-9.]*\|tc_counters.: ([0-9, ]*)\|idx_mismatch.: [0-9]*\|band_check.*" | paste - - - - 
grep -i "exc\|error" gpurun_out/tc_debug.log | head -5
echo "== pytest vq"; timeout 900 python -m pytest tests/test_vq_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_vq.log
echo "== sweep"; timeout 900 python scripts/bench_sweep_full.py --iters 2 ${SWEEP_ARGS} 2>&1 | grep "^{" | cut -c1-210
echo "== sweep DVQ_TC_ST=1"; DVQ_TC_ST=1 timeout 900 python scripts/bench_sweep_full.py --iters 2 ${SWEEP_AB_ARGS:-$SWEEP_ARGS} 2>&1 | grep "^{" | cut -c1-210
echo "== sweep DVQ_TC_CE=1"; DVQ_TC_CE=1 timeout 900 python scripts/bench_sweep_full.py --iters 2 ${SWEEP_AB_ARGS:-$SWEEP_ARGS} 2>&1 | grep "^{" | cut -c1-210
echo "== bench"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | grep -o "\"value\": [0-9.]*\|ms_per_step\": [0-9.]*\|kernel_ms\": [0-9.]*\|refine_ms\": [0-9.]*\|frac\": [0-9.]*" | paste - - - - - - - -
