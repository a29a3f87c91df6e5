#!/bin/bash
# ring / staging / A-image layouts of the CTA-pair kernel (DVQ_TC_LAYOUT=<A images><z staging slots><ring slots>, up to 8 slots)
export ABNAME=pairlay
N=${N:-4194304}
specs=""
for kd in 16384,128 4096,128 16384,256 4096,256 16384,512 4096,512; do
  K=${kd%%,*}; D=${kd##*,}
  specs="$specs base:N=$N,K=$K,D=$D,STEPS=4"
  for lay in ${LAYS:-213 214 215 113 115 123 124 118}; do specs="$specs base:DVQ_TC_LAYOUT=$lay,N=$N,K=$K,D=$D,STEPS=4"; done
done
bash scripts/gpu_ab.sh "$specs" > /dev/null
python - <<PY
import json
for l in open("gpurun_out/ab_pairlay.jsonl"):
    d = json.loads(l)
    if "kernel_ms" in d: print("%-58s kernel %8.3f refine %7.3f mism %d" % (d["tag"], d["kernel_ms"], d["refine_ms"], d["idx_mismatch_vs_simt"]))
    else: print(d)
PY
