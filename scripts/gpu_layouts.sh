#!/bin/bash
# smem_layout override experiment (DVQ_TC_LAYOUT = digits A-images / z staging slots / ring slots) on a few config-4 shapes
for lay in ${LAYOUTS:-0 113 112 0 113}; do
  echo "== layout $lay"
  DVQ_TC_LAYOUT=$lay python scripts/bench_sweep_full.py --shapes ${SHAPES:-512:128,256:128,1024:128} 2>/dev/null | grep -o "\"K\": [0-9]*, \"D\": [0-9]*\|filter_ms\": [0-9.]*\|\"ms\": [0-9.]*" | paste - - -
done
