#!/bin/bash
# resident CTA-pair layout: parity at an odd tile count, then timing at full size, against the single-CTA kernel
export ABNAME=rpair
specs=""
for kd in ${SHAPES:-512,128 1024,128 1024,64 2048,64 256,256 768,128}; do
  K=${kd%%,*}; D=${kd##*,}
  specs="$specs base:DVQ_TC_PAIR=1,N=300109,K=$K,D=$D,STEPS=3 base:DVQ_TC_PAIR=1,N=16777216,K=$K,D=$D,STEPS=5 base:DVQ_TC_PAIR=0,N=16777216,K=$K,D=$D,STEPS=5"
done
bash scripts/gpu_ab.sh "$specs" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-62s kernel %8.3f refine %7.3f step %8.3f mism %d zq %s cnt %s' % (d['tag'], d['kernel_ms'], d['refine_ms'], d['step_ms'], d['idx_mismatch_vs_simt'], d['zq_equal'], d['counters']))
"
tail -3 gpurun_out/ab_rpair.err
