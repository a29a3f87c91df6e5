#!/bin/bash
# A/B run of kernel-variant builds (libdvq_sm100_<tag>.so) at config 2: scripts/ab_config2.py per variant.
# usage: bash scripts/gpu_ab.sh "tag[:ENV=val[,ENV=val]] ..."   ("base" = the default build)
mkdir -p gpurun_out
OUT=gpurun_out/ab_${ABNAME:-run}.jsonl
: > $OUT
for spec in $1; do
  tag=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  lib=$PWD/d-vqvae_b200/dvq/libdvq_sm100_${tag}.so
  [ "$tag" == "base" ] && lib=$PWD/d-vqvae_b200/dvq/libdvq_sm100.so
  env $envs TAG=$spec DVQ_LIB=$lib timeout 120 python scripts/ab_config2.py >> $OUT 2>> gpurun_out/ab_${ABNAME:-run}.err || echo "{\"tag\": \"$spec\", \"failed\": true}" >> $OUT
done
cat $OUT | cut -c1-400
