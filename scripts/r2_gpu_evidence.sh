#!/bin/bash
# Round-2 single-GPU evidence (run under gpurun): GPU tests, the driver's bench invocation, the reference arm,
# the ncu launch list and the full captures of the dominant kernels.  Everything lands in gpurun_out/<tag>_*.
# Usage: scripts/r2_gpu_evidence.sh <tag> [notests] [noref]
tag=${1:-r2}; shift
OUT=gpurun_out
mkdir -p $OUT
if [[ " $* " != *" notests "* ]]; then
  python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${tag}_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/${tag}_gputests.log
fi
python bench.py --steps 20 --warmup 5 > $OUT/${tag}_bench_1gpu.json 2> $OUT/${tag}_bench_1gpu.err; echo "bench rc=$?"
python - $OUT/${tag}_bench_1gpu.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.4f e2e %.4f launches %d clocks %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]))
for k in d["kernels"]:
    print("   %-36s %.4f" % (k["name"], k["ms"]))
r = d["roofline"]
print("roofline", {k: r[k] for k in ("kernel", "achieved", "frac", "traffic", "kernel_ms")}, r["issue"])
print("cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
PY
if [[ " $* " != *" noref "* ]]; then
  python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${tag}_bench_reference.json 2> $OUT/${tag}_bench_reference.err; echo "reference rc=$?"
  python bench.py --workload batch --steps 20 --warmup 5 --no-cpu > $OUT/${tag}_bench_batch.json 2> $OUT/${tag}_bench_batch.err; echo "batch rc=$?"
fi
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# full captures: the ray kernel, then the resampling search, the table product and the normalising pass
ncu --set full --clock-control none --import-source on -k regex:k_raycast_dir -s 6 -c 1 -f -o $OUT/prof_${tag}_dir \
  python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${tag}_ncu_dir.log 2>&1; echo "ncu dir rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_resample_motion|k_weight_steps|k_exact_pass|k_exact_emit|k_tile_sums|k_dir_gather|k_sort' -s 40 -c 12 -f -o $OUT/prof_${tag}_small \
  python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${tag}_ncu_small.log 2>&1; echo "ncu small rc=$?"
ls -la $OUT | tail -20
