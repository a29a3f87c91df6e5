#!/bin/bash
# pair kernel: role wait accounting (instrumented build) and ring-layout experiments at one shape
mkdir -p gpurun_out
K=${K:-16384}; D=${D:-128}; N=${N:-2097152}
for lay in ${LAYOUTS:-0}; do
  echo "== pair layout $lay"
  DVQ_TC_LAYOUT=$lay N=$N K=$K D=$D DVQ_LIB=$PWD/d-vqvae_b200/dvq/libdvq_sm100_stats.so DVQ_TC_STATS_PRINT=1 ITERS=2 timeout 300 python scripts/tc_profile.py 2>&1 | grep -v "trace\]" | tail -19
  DVQ_TC_LAYOUT=$lay N=$N K=$K D=$D STEPS=5 TAG=pair_lay$lay timeout 300 python scripts/ab_config2.py | cut -c1-200
  for t in ${TAGS}; do
    DVQ_LIB=$PWD/d-vqvae_b200/dvq/libdvq_sm100_$t.so DVQ_TC_LAYOUT=$lay N=$N K=$K D=$D STEPS=5 TAG=pair_${t}_lay$lay timeout 300 python scripts/ab_config2.py | cut -c1-200
  done
done
echo "== solo"
DVQ_TC_PAIR=0 N=$N K=$K D=$D STEPS=5 TAG=solo timeout 300 python scripts/ab_config2.py | cut -c1-200
