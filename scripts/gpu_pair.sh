#!/bin/bash
# CTA-pair kernel (cta_group::2) vs the single-CTA kernel: parity against the FP32 kernel at an odd tile count, then timing at full size.
# usage: bash scripts/gpu_pair.sh [quick|full]
export ABNAME=pair_${1:-quick}
NQ=300109
if [ "${1:-quick}" == "quick" ]; then
  specs=""
  for kd in 1024,64 4096,128 16384,128 2048,256 4096,512 544,128 8192,64; do
    K=${kd%%,*}; D=${kd##*,}
    specs="$specs base:N=$NQ,K=$K,D=$D,STEPS=5 base:DVQ_TC_PAIR=0,N=$NQ,K=$K,D=$D,STEPS=5"
  done
else
  specs=""
  for kd in ${SHAPES:-16384,128 4096,128 4096,256 4096,512 16384,512 8192,64 1024,128 16384,64}; do
    K=${kd%%,*}; D=${kd##*,}
    specs="$specs base:N=16777216,K=$K,D=$D,STEPS=5 base:DVQ_TC_PAIR=0,N=16777216,K=$K,D=$D,STEPS=5"
  done
fi
bash scripts/gpu_ab.sh "$specs"
tail -5 gpurun_out/ab_${ABNAME}.err 2>/dev/null
