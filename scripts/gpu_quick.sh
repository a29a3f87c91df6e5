timeout 240 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo rc=$?
cut -c1-330 gpurun_out/tc_debug.log | grep -o "^([0-9, a-z_']*)\|tc_ms.: [0-9.]*\|tc_counters.: ([0-9, ]*)\|idx_mismatch.: [0-9]*\|band_check.*" | paste - - - - 
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log | grep -o "\"value\": [0-9.]*\|ms_per_step\": [0-9.]*\|kernel_ms\": [0-9.]*\|refine_ms\": [0-9.]*\|frac\": [0-9.]*" | head -5
