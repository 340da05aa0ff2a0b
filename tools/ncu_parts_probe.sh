#!/bin/bash
# GPU box (one GPU): ncu --set full captures of kernels of the domain-partitioned step, run with 8 virtual parts
# (NB_BH_PARTS=8: the same kernels a rank of an 8-GPU job runs, on 1/8 of the bodies each).  Usage: tools/ncu_parts_probe.sh <tag> <kernel regex>...
T=$1; shift
for K in "$@"; do
  NB_BH_PARTS=8 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 27 -c 1 -f -o gpurun_out/${T}_$K python bench.py --workload c5 --steps 3 --no-parity --no-cpu-baseline > /dev/null 2> gpurun_out/${T}_$K.err
  tail -2 gpurun_out/${T}_$K.err
done
