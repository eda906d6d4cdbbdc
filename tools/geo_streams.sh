set -x
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-e2e"
for sc in 1 2 8; do
  for gs in 1 4; do
    STB200_GEOMETRY_STREAMS=$gs $B --scenes $sc > gpurun_out/gs_${sc}_${gs}.json 2> gpurun_out/gs_${sc}_${gs}.err
    python - gpurun_out/gs_${sc}_${gs}.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['ms_per_step'], d['value'], {k:v for k,v in d.items() if 'geometry' in k})
P
  done
done
