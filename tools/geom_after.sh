# experiment: when inside the step the next batch's geometry is submitted (bench.py, STB200_BENCH_GEOM_AFTER_LEVEL)
for k in ${@:--1 0 1 2}; do
  STB200_BENCH_GEOM_AFTER_LEVEL=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda $EXTRA > gpurun_out/ga_$k.json 2>gpurun_out/ga_$k.err
  python - gpurun_out/ga_$k.json $k <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print("geometry after level", sys.argv[2], round(d["ms_per_step"],2), "attention kernel sum", round(d["roofline"]["path"]["kernel_ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
P
done
