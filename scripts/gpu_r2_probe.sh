#!/bin/bash
# round 2, first GPU call: pipe-rate microbenchmark, role wait accounting + timeline of the config-2 kernel
# (instrumented build), and the unchanged round-1 bench as the baseline of this round.
mkdir -p gpurun_out
timeout 120 scripts/ubench/pipes > gpurun_out/r02_pipes.txt 2>&1
DVQ_LIB=$PWD/d-vqvae_b200/dvq/libdvq_sm100_stats.so DVQ_TC_STATS_PRINT=1 ITERS=3 timeout 300 python scripts/tc_profile.py > gpurun_out/r02_stats.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err
tail -c 600 gpurun_out/r02_bench0.json
