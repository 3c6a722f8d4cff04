set -x
for rep in 1 2 3; do
for bar in "" 1; do
for n in 524288; do
CHECK_BARRIER_AFTER_INIT=$bar MCL_PACKED_PEERS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$rep scripts/check_sharded_equals_single.py --particles-per-gpu $n --updates 8 2>/dev/null | grep '^{' | sed "s/^/{\"barrier_after_init\": \"$bar\", \"packed_peers\": 1, \"rep\": $rep, \"res\": /; s/$/}/" >> gpurun_out/r2_race_probe.jsonl
done; done; done
cat gpurun_out/r2_race_probe.jsonl
