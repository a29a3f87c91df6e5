#!/bin/bash
# Round-end single-GPU evidence: smoke, all GPU tests, bench line, ncu launch list + full captures, config-5 driver on one GPU.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-300
echo "== grasp dist (1 GPU, 1250 objects x 100 grasps)"; N_OBJ=1250 timeout 600 python scripts/bench_grasp_dist.py 2>&1 | tail -2 | cut -c1-400
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full: main kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01d_tc \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r01d_tc.log 2>&1; tail -1 gpurun_out/ncu_full_r01d_tc.log
echo "== ncu full: per-row refine (config 2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_refine_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01d_refine \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r01d_refine.log 2>&1; tail -1 gpurun_out/ncu_full_r01d_refine.log
echo "== ncu full: binned refine pairs kernel (K=4096, D=128, N=1M)"
N=1048576 K=4096 D=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:refine_pairs_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01d_pairs \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r01d_pairs.log 2>&1; tail -1 gpurun_out/ncu_full_r01d_pairs.log
