#!/bin/bash
# Development aid: the default bench line for several builds of the library (scripts/build_variant.sh), one after the other
# on the same box.  Usage: scripts/bench_variants.sh <tag> <name> [<name> ...]   ("default" = the in-tree library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset MCL_B200_LIB; else export MCL_B200_LIB=$PWD/build/variants/lib$v.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$v" gpurun_out/${tag}_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = {x["name"]: x["ms"] for x in d.get("kernels", [])}
    print("%-10s ms/step %.4f e2e %.4f ray %.4f route/motion %.4f weight %.4f sum_kernels %.4f" % (
        sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], k.get("k_raycast_dir", 0), k.get("k_resample_motion", 0),
        k.get("k_weight_steps", 0), sum(k.values())))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
