#!/bin/bash
# four-stage pair kernel (DVQ_TC_PAIR4=1): parity at an odd tile count, then timing against the two-stage pair kernel and the single-CTA kernel
export ABNAME=pair4
N=${N:-4194304}
specs=""
for kd in ${QSHAPES:-1024,64 4096,128 2048,256 1024,512 544,128 8192,64 16384,128}; do
  K=${kd%%,*}; D=${kd##*,}
  specs="$specs base:DVQ_TC_PAIR=1,DVQ_TC_PAIR4=1,N=300109,K=$K,D=$D,STEPS=3"
done
for kd in ${SHAPES:-16384,64 4096,64 2048,64 16384,128 4096,128 4096,256 16384,512}; do
  K=${kd%%,*}; D=${kd##*,}
  specs="$specs base:DVQ_TC_PAIR=1,DVQ_TC_PAIR4=1,N=$N,K=$K,D=$D,STEPS=5 base:DVQ_TC_PAIR=1,N=$N,K=$K,D=$D,STEPS=5 base:DVQ_TC_PAIR=0,N=$N,K=$K,D=$D,STEPS=5"
done
bash scripts/gpu_ab.sh "$specs" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-72s kernel %8.3f refine %7.3f step %8.3f mism %d zq %s cnt %s' % (d['tag'], d['kernel_ms'], d['refine_ms'], d['step_ms'], d['idx_mismatch_vs_simt'], d['zq_equal'], d['counters']))
"
tail -5 gpurun_out/ab_pair4.err
