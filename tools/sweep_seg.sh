#!/bin/bash
# Sweep launch knobs of the segment kernels and print per-kernel ms per step (overlapped bench step, then solo).
for t in 0 256 128 384; do
  if [ "$t" = "0" ]; then unset STB200_SEGRED_THREADS; else export STB200_SEGRED_THREADS=$t; fi
  for mode in "" "--no-overlap"; do
  echo "== segred_threads=$t $mode"
  python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 $mode | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],2))
k=d['roofline']['per_kernel_ms_per_step']
print({a:b for a,b in k.items() if a.startswith('seg_reduce')})
"
  done
done
