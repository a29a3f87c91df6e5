#!/bin/bash
# Round-2 single-GPU evidence: smoke, all GPU tests, both bench arms, full-size sweep, grasp driver, ncu launch list of bench.py,
# full ncu captures of the config-2 filter kernel and of the pipelined PointNet trunk.  scripts/summarize_profiles.py r02 ... turns
# the outputs into profiles/r02_*.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench_n1.err | tail -1 > gpurun_out/r02_bench_n1.json; cut -c1-300 gpurun_out/r02_bench_n1.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_n1.json; cut -c1-300 gpurun_out/r02_bench_reference_n1.json
echo "== sweep"; timeout 900 python scripts/bench_sweep_full.py 2>/dev/null | grep -c "^{"; cp gpurun_out/sweep_full.json gpurun_out/r02_sweep_full_1gpu.json
echo "== pointnet"; python scripts/bench_pointnet.py 2>/dev/null | tail -1 | tee gpurun_out/r02_pointnet.json
echo "== grasp dist (1 GPU, 1250 objects x 100 grasps)"; N_OBJ=1250 timeout 600 python scripts/bench_grasp_dist.py 2>&1 | tail -1 | tee gpurun_out/r02_grasp_dist_1gpu.json | cut -c1-300
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu launches: PointNet encoder"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pointnet|stn_head|decode' -c 60 --csv --log-file gpurun_out/launches_pointnet.csv \
    python scripts/bench_pointnet.py > gpurun_out/ncu_pointnet_launches.log 2>&1
echo "== ncu full: config-2 filter kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_tc \
    python scripts/tc_profile.py > gpurun_out/ncu_full_r02_tc.log 2>&1; tail -1 gpurun_out/ncu_full_r02_tc.log
echo "== ncu full: pipelined PointNet trunk"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointnet_trunk_tc2 -s 2 -c 1 -f -o gpurun_out/prof_r02_pointnet \
    python scripts/pn_profile.py > gpurun_out/ncu_full_r02_pointnet.log 2>&1; tail -1 gpurun_out/ncu_full_r02_pointnet.log
