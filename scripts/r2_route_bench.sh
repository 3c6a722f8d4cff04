#!/bin/bash
# bench.py at N GPUs with both routings of the resampling draws (no parity run).  Usage: scripts/r2_route_bench.sh N [tag]
N=${1:-4}; tag=${2:-r2route}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
for route in two-hop one-hop; do
  $TR --master-port 29554 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --shard-route $route > $OUT/${tag}_bench_${N}gpu_$route.json 2> $OUT/${tag}_bench_${N}gpu_$route.err; echo "bench $route rc=$?"
done
python - $OUT/${tag}_bench_${N}gpu_two-hop.json $OUT/${tag}_bench_${N}gpu_one-hop.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"])
        print("   " + "  ".join("%s %.3f" % (k["name"].replace("k_", ""), k["ms"]) for k in d["kernels"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
