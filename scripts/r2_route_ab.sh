#!/bin/bash
# Two-hop request routing against one-hop (every rank tests all draws) on N real GPUs: bitwise parity with one GPU, then
# the bench in both modes.  Usage (under gpurun --gpus N): scripts/r2_route_ab.sh N [tag]
N=${1:-2}; tag=${2:-r2route}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
$TR --master-port 29551 scripts/check_sharded_equals_single.py --particles-per-gpu 1048576 --updates 8 --degenerate 2>$OUT/${tag}_par_${N}.err | grep '^{' > $OUT/${tag}_parity_${N}gpu.jsonl
$TR --master-port 29552 scripts/check_sharded_equals_single.py --particles-per-gpu 262144 --updates 8 --exchange nccl 2>>$OUT/${tag}_par_${N}.err | grep '^{' >> $OUT/${tag}_parity_${N}gpu.jsonl
cat $OUT/${tag}_parity_${N}gpu.jsonl
for route in two-hop one-hop; do
  $TR --master-port 29554 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --shard-route $route > $OUT/${tag}_bench_${N}gpu_$route.json 2> $OUT/${tag}_bench_${N}gpu_$route.err; echo "bench $route rc=$?"
done
python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
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
