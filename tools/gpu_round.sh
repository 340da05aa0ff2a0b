#!/bin/bash
# GPU box: the round-trip used while tuning the single-GPU Barnes-Hut step -- GPU tests, per-phase probes at c4 / c5 (disk and
# Plummer), and an ncu launch list of the c5 step with warm caches.  Usage: tools/gpu_round.sh <tag> [skip-tests]
T=$1
O=gpurun_out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/${T}_pytest.txt; tail -4 $O/${T}_pytest.txt
fi
(python tools/bh_phase_probe.py; N=4194304 THETA=0.75 python tools/bh_phase_probe.py; N=4194304 THETA=0.75 GEN=plummer python tools/bh_phase_probe.py; NB_SORT_LEVELS=12 N=4194304 THETA=0.75 python tools/bh_phase_probe.py; NB_SORT_LEVELS=12 python tools/bh_phase_probe.py) > $O/${T}_probe.jsonl 2> $O/${T}_probe.err
python - <<P
import json
for l in open("$O/${T}_probe.jsonl"):
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print(d["n"], d["gen"], d["forced_levels"], round(d["ms_per_step_back_to_back"], 4), {k: round(v, 4) for k, v in d["phases_ms"].items()})
P
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 260 --csv --log-file $O/${T}_launches_c5.csv python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2> $O/${T}.err; tail -3 $O/${T}.err
