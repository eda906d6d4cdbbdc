# diagnostic: where does the multi-rank weak-scaling loss come from (data-dependent straggler / exchange / host)?
N=${1:-4}
PORT=29600
run() { tag=$1; shift; PORT=$((PORT+1))
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-e2e > gpurun_out/diag_$tag.json 2> gpurun_out/diag_$tag.err
  python - gpurun_out/diag_$tag.json $tag <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], round(d['ms_per_step'],2), round(d['value']/1e6,2))
P
}
nproc
run default A=1
run same_scenes STB200_BENCH_SAME_SCENES=1
run no_allreduce STB200_BENCH_NO_ALLREDUCE=1
run same_no_allreduce STB200_BENCH_SAME_SCENES=1 STB200_BENCH_NO_ALLREDUCE=1
