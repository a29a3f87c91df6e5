#!/bin/bash
# second-session check: tc vs FP32 kernel on many shapes, VQ GPU tests, bench (binned vs per-row refine), full-size sweep subset
mkdir -p gpurun_out
# canary: one small call first; a launch failure here ends the script (a dead context costs GPU minutes)
N=65536 ITERS=2 timeout 90 python scripts/tc_profile.py 2>&1 | tail -2 | tee gpurun_out/canary.log; grep -q "^loss" gpurun_out/canary.log || { echo CANARY FAILED; exit 1; }
timeout 300 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo rc=$?
cut -c1-330 gpurun_out/tc_debug.log | grep -o "^([0-9, a-z_']*)\|tc_ms.: [0-9.]*\|tc_counters.: ([0-9, ]*)\|idx_mismatch.: [0-9]*\|band_check.*" | paste - - - - 
grep -i "exc\|error" gpurun_out/tc_debug.log | head -5
echo "== pytest vq"; timeout 900 python -m pytest tests/test_vq_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_vq.log
echo "== bench binned"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | tee gpurun_out/bench_binned.log | grep -o "\"value\": [0-9.]*\|ms_per_step\": [0-9.]*\|kernel_ms\": [0-9.]*\|refine_ms\": [0-9.]*\|frac\": [0-9.]*" | head -5
echo "== bench per-row"; DVQ_REFINE_PER_ROW=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | tee gpurun_out/bench_perrow.log | grep -o "\"value\": [0-9.]*\|ms_per_step\": [0-9.]*\|kernel_ms\": [0-9.]*\|refine_ms\": [0-9.]*\|frac\": [0-9.]*" | head -5
echo "== sweep"; timeout 600 python scripts/bench_sweep_full.py --iters 2 ${SWEEP_ARGS} 2>&1 | tail -30 | cut -c1-420
if [ -n "$SWEEP_AB" ]; then echo "== sweep per-row"; DVQ_REFINE_PER_ROW=1 timeout 600 python scripts/bench_sweep_full.py --iters 2 ${SWEEP_ARGS} 2>&1 | tail -30 | cut -c1-200; fi
