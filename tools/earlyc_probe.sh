#!/bin/bash
# GPU box: A/B of the walk's child-index fetch (NB_BH_EARLYC = 0 conditional load, 1 early register load, 2 L1 prefetch)
for E in 0 1 2; do NB_BH_EARLYC=$E python tools/bh_phase_probe.py; NB_BH_EARLYC=$E N=4194304 THETA=0.75 python tools/bh_phase_probe.py; done > gpurun_out/$1_earlyc.jsonl 2>&1
python - <<P
import json
for l in open("gpurun_out/$1_earlyc.jsonl"):
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print(d["n"], round(d["ms_per_step_back_to_back"], 4), round(d["phases_ms"]["force"], 4))
P
