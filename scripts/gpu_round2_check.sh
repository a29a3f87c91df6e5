#!/bin/bash
# full GPU regression of round 2: every -m gpu test, smoke(), the bench line (both arms).
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 1500 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
