#!/bin/bash
# GPU box with $1 GPUs: the default bench line at N = $1 (c3 headline + c5 Barnes-Hut sub-record + parity check), and, at 8, the
# sharded nbx3 workload and the multi-GPU parity checks.
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29551 bench.py --gpus $N > $O/r02_bench_g$N.json 2> $O/r02_bench_g$N.err; tail -c 300 $O/r02_bench_g$N.err
python - <<P
import json
d = json.load(open("$O/r02_bench_g$N.json"))
print("c3", d["ms_per_step"], d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
b = d["bh"]["c5"]
print("c5", b["ms_per_step"], b["steps_per_s"], b["phases_ms"], "e2e", b["e2e"]["ms_per_step"], b.get("ordering_point_ms"), b.get("walk_imbalance_max_over_mean"))
for r in b.get("per_rank", []): print(r["rank"], r["part_bodies"], r["force"], r["sort"], r["build"], r["xrank_boxes_bodies_trees_walks"], r["com_merge_top"])
print(d.get("parity_check"))
P
if [ "$N" = "8" ]; then
  $TR --master-port 29552 bench.py --gpus $N --workload c3n --steps 5 --no-parity > $O/r02_bench_c3n_g$N.json 2> $O/r02_bench_c3n_g$N.err; cut -c1-200 $O/r02_bench_c3n_g$N.json
  NB_DIST_BIG=262144 NB_DIST_BH=1048576 $TR --master-port 29553 tools/dist_check.py > $O/r02_dist_check_8gpu.jsonl 2> $O/r02_dist_check_8gpu.err; grep -c '"pass": true' $O/r02_dist_check_8gpu.jsonl; grep '"pass": false' $O/r02_dist_check_8gpu.jsonl; grep time_ $O/r02_dist_check_8gpu.jsonl
fi
