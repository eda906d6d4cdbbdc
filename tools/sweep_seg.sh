#!/bin/bash
# Development aid: A/B the launch knobs of the hot kernels on the bench step (overlapped, device-resident).
#   tools/sweep_seg.sh "STB200_TG_WAVES=2" "STB200_SEGRED_THREADS=512" ...
# Each argument is a space-separated list of VAR=value settings for one run; a baseline run comes first and last.
run() {
  env $1 python bench.py --no-e2e --no-cpu-baseline --steps 12 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['per_kernel_ms_per_step']
print('%-40s %7.2f ms/step   ' % ('$1' or 'baseline', d['ms_per_step']), {a: round(b, 2) for a, b in k.items() if a.startswith(('seg_', 'table'))})
"
}
run ""
for cfg in "$@"; do run "$cfg"; done
run ""
