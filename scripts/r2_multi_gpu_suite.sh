#!/bin/bash
# Round-2 multi-GPU evidence (run under gpurun --gpus N): bitwise parity of the sharded filter against one GPU,
# the C++ / pytest multi-GPU tests, and the bench at N (and N/2) ranks.  Usage: r2_multi_gpu_suite.sh N
N=${1:-2}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node $N --master-port 29551 scripts/check_sharded_equals_single.py --particles-per-gpu 262144 --updates 8 --degenerate 2>$OUT/r2_par_${N}a.err | grep '^{' > $OUT/r2_par_${N}gpu.jsonl
$TR --nproc-per-node $N --master-port 29552 scripts/check_sharded_equals_single.py --particles-per-gpu 1048576 --updates 12 2>$OUT/r2_par_${N}b.err | grep '^{' >> $OUT/r2_par_${N}gpu.jsonl
$TR --nproc-per-node $N --master-port 29553 scripts/check_sharded_equals_single.py --particles-per-gpu 262144 --updates 8 --exchange nccl 2>$OUT/r2_par_${N}c.err | grep '^{' >> $OUT/r2_par_${N}gpu.jsonl
cat $OUT/r2_par_${N}gpu.jsonl
$TR --nproc-per-node $N --master-port 29554 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2_bench_${N}gpu.json 2> $OUT/r2_bench_${N}gpu.err; echo "bench $N rc=$?"
if [ "$N" -ge 4 ]; then
  H=$((N/2))
  $TR --nproc-per-node $H --master-port 29555 bench.py --gpus $H --steps 20 --warmup 5 > $OUT/r2_bench_${H}gpu.json 2> $OUT/r2_bench_${H}gpu.err; echo "bench $H rc=$?"
fi
$TR --nproc-per-node $N --master-port 29556 bench.py --gpus $N --steps 20 --warmup 5 --shard-exchange nccl > $OUT/r2_bench_${N}gpu_nccl.json 2> $OUT/r2_bench_${N}gpu_nccl.err; echo "bench nccl rc=$?"
python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*gpu*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], d["config"]["sharding"][-20:])
    except Exception as e:
        print(f, "unreadable", e)
PY
