#!/bin/bash
# ncu full capture of one kernel (regex in $KREGEX) + a launch list of the bench command.
mkdir -p gpurun_out
KR=${KREGEX:-vq_tc_kernel}   # (not K: tc_profile.py reads N, K, D from the environment)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s ${SKIP:-2} -c 1 -f -o gpurun_out/prof_${TAG:-tc} \
    python scripts/tc_profile.py > gpurun_out/ncu_full_${TAG:-tc}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG:-tc}.log
