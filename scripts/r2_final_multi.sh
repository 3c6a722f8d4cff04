#!/bin/bash
# Final round-2 multi-GPU record (run under gpurun --gpus N): bitwise parity of the sharded filter against one GPU (two-hop
# routing, in-kernel exchanges), bench at N ranks with 1 M and 2 M particles per rank, the batch workload filter-sharded
# over N, and BASELINE config 5 (global initialisation).  Usage: scripts/r2_final_multi.sh N
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
$TR --master-port 29551 scripts/check_sharded_equals_single.py --particles-per-gpu 1048576 --updates 8 --degenerate 2>$OUT/r2f_par_${N}.err | grep '^{' > $OUT/r2f_parity_${N}gpu.jsonl; echo "parity rc=$?"
cat $OUT/r2f_parity_${N}gpu.jsonl
$TR --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2f_bench_${N}gpu.json 2> $OUT/r2f_bench_${N}gpu.err; echo "bench rc=$?"
$TR --master-port 29553 bench.py --gpus $N --steps 20 --warmup 5 --particles 2097152 > $OUT/r2f_bench_${N}gpu_2M.json 2> $OUT/r2f_bench_${N}gpu_2M.err; echo "bench 2M rc=$?"
$TR --master-port 29554 bench.py --gpus $N --steps 20 --warmup 5 --shard-route one-hop > $OUT/r2f_bench_${N}gpu_onehop.json 2> $OUT/r2f_bench_${N}gpu_onehop.err; echo "bench one-hop rc=$?"
$TR --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 --workload batch > $OUT/r2f_bench_${N}gpu_batch.json 2> $OUT/r2f_bench_${N}gpu_batch.err; echo "bench batch rc=$?"
$TR --master-port 29556 scripts/run_config5_sharded.py 2> $OUT/r2f_config5_${N}gpu.err | grep '^{' > $OUT/r2f_config5_${N}gpu.json; echo "config5 rc=$?"
python - $OUT/r2f_bench_${N}gpu.json $OUT/r2f_bench_${N}gpu_2M.json $OUT/r2f_bench_${N}gpu_onehop.json $OUT/r2f_bench_${N}gpu_batch.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "value %.3e" % d["value"])
        print("   " + "  ".join("%s %.3f" % (k["name"].replace("k_", ""), k["ms"]) for k in d["kernels"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
cut -c1-600 $OUT/r2f_config5_${N}gpu.json
