#!/bin/bash
# Sweep table-gradient launch knobs (heads per CTA, waves) and print per-kernel ms per step, solo and overlapped.
for cfg in "3 1" "2 1" "1 1" "1 2" "2 2"; do
  set -- $cfg
  export STB200_TG_HEADS=$1 STB200_TG_WAVES=$2
  for mode in "--no-overlap" ""; do
  echo "== tg_heads=$1 tg_waves=$2 $mode"
  python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 $mode | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],2))
k=d['roofline']['per_kernel_ms_per_step']
print({a:b for a,b in k.items() if a.startswith('table')})
"
  done
done
