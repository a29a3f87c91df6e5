#!/bin/bash
# final single-GPU evidence of the round: tests, bench, full-size sweep, ncu launch list and full capture
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
echo "== sweep"; timeout 900 python scripts/bench_sweep_full.py --iters 2 2>&1 | grep "^{" | cut -c1-140
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: main kernel (config 2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01f_tc \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r01f_tc.log 2>&1; tail -1 gpurun_out/ncu_full_r01f_tc.log
echo "== ncu full: main kernel (K=16384, e_dim 128, N=1M)"
N=1048576 K=16384 D=128 ITERS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01f_tc_k16k_d128 \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r01f_k16k.log 2>&1; tail -1 gpurun_out/ncu_full_r01f_k16k.log
