#!/bin/bash
# Round-2 final single-GPU evidence (after the CTA-pair kernel): smoke, all GPU tests, both bench arms, full-size sweep, PointNet,
# grasp driver, ncu launch lists, full ncu captures of the config-2 filter kernel, of the CTA-pair kernel (K = 16384 at e_dim 128
# and 512) and of the single-CTA kernel at the same e_dim 128 shape.  scripts/summarize_profiles.py turns the outputs into profiles/.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench_n1.err | tail -1 > gpurun_out/r02_bench_n1.json; cut -c1-300 gpurun_out/r02_bench_n1.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_n1.json; cut -c1-300 gpurun_out/r02_bench_reference_n1.json
echo "== sweep"; timeout 900 python scripts/bench_sweep_full.py 2>/dev/null | grep -c "^{"; cp gpurun_out/sweep_full.json gpurun_out/r02_sweep_full_1gpu_pair.json
echo "== pointnet"; python scripts/bench_pointnet.py 2>/dev/null | tail -1 | tee gpurun_out/r02_pointnet.json
echo "== grasp dist (1 GPU, 1250 objects x 100 grasps)"; N_OBJ=1250 timeout 600 python scripts/bench_grasp_dist.py 2>&1 | tail -1 | tee gpurun_out/r02_grasp_dist_1gpu.json | cut -c1-300
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu launches: PointNet encoder"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pointnet|stn_head|decode' -c 60 --csv --log-file gpurun_out/launches_pointnet.csv \
    python scripts/bench_pointnet.py > gpurun_out/ncu_pointnet_launches.log 2>&1
echo "== ncu full: config-2 filter kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_tc \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r02_tc.log 2>&1; tail -1 gpurun_out/ncu_full_r02_tc.log
echo "== ncu full: CTA-pair kernel, K = 16384, e_dim 128 (N = 1M)"
N=1048576 K=16384 D=128 ITERS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_pair_k16k_d128 \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r02_pair_d128.log 2>&1; tail -1 gpurun_out/ncu_full_r02_pair_d128.log
echo "== ncu full: single-CTA kernel, same shape (DVQ_TC_PAIR=0)"
DVQ_TC_PAIR=0 N=1048576 K=16384 D=128 ITERS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_solo_k16k_d128 \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r02_solo_d128.log 2>&1; tail -1 gpurun_out/ncu_full_r02_solo_d128.log
echo "== ncu full: CTA-pair kernel, K = 16384, e_dim 512 (N = 512k)"
N=524288 K=16384 D=512 ITERS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_pair_k16k_d512 \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r02_pair_d512.log 2>&1; tail -1 gpurun_out/ncu_full_r02_pair_d512.log
ls -la gpurun_out/*.ncu-rep
