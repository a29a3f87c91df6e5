#!/bin/bash
# two-sub-chunk filter (DVQ_TC_ST=1) with the list-mode early-out vs the default, large codebooks
export ABNAME=st
N=${N:-4194304}
specs=""
for kd in ${SHAPES:-4096,64 8192,64 16384,64 2048,64 16384,128 4096,128 16384,256}; do
  K=${kd%%,*}; D=${kd##*,}
  specs="$specs base:N=$N,K=$K,D=$D,STEPS=5 base:DVQ_TC_ST=1,N=$N,K=$K,D=$D,STEPS=5"
done
bash scripts/gpu_ab.sh "$specs" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-60s kernel %8.3f refine %7.3f step %8.3f mism %d zq %s' % (d['tag'], d['kernel_ms'], d['refine_ms'], d['step_ms'], d['idx_mismatch_vs_simt'], d['zq_equal']))
"
tail -3 gpurun_out/ab_st.err
