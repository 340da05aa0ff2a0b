#!/bin/bash
# Runs on the GPU box (gpurun, ONE GPU): the bench lines, the ncu launch lists, the --set full captures of the dominant
# kernels and the compute-sanitizer passes; everything lands in gpurun_out/ and is summarised into profiles/ afterwards
# (tools/ncu_summary.py, tools/sass_summary.py).
set -x
O=gpurun_out
python bench.py > $O/r02_bench_g1.json 2> $O/r02_bench_g1.err; tail -c 300 $O/r02_bench_g1.err
python bench.py --workload c3n --steps 5 --no-parity > $O/r02_bench_c3n_g1.json 2> $O/r02_bench_c3n_g1.err
python bench.py --workload c2 --steps 20 --no-parity > $O/r02_bench_c2_g1.json 2> $O/r02_bench_c2_g1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2> $O/r02_launches_default.err
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 260 --csv --log-file $O/r02_launches_c5_warm.csv python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2> $O/r02_launches_c5.err
ncu --set full --clock-control none --import-source on -k regex:allpairs_fast --launch-skip 3 -c 1 -f -o $O/r02_allpairs_c3 python bench.py --steps 1 --warmup 3 --no-bh --no-parity --no-cpu-baseline > /dev/null 2> $O/r02_ncu_allpairs.err
ncu --set full --clock-control none --import-source on -k regex:bh_traverse_fast --launch-skip 4 -c 1 -f -o $O/r02_walk_c4 python bench.py --workload c4 --steps 3 --no-parity --no-cpu-baseline > /dev/null 2> $O/r02_ncu_c4.err
ncu --set full --clock-control none --import-source on -k regex:bh_traverse_fast --launch-skip 4 -c 1 -f -o $O/r02_walk_c5 python bench.py --workload c5 --steps 3 --no-parity --no-cpu-baseline > /dev/null 2> $O/r02_ncu_c5.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > $O/r02_sanitizer_$tool.txt 2>&1
  tail -3 $O/r02_sanitizer_$tool.txt
done
python tools/mode_costs.py > $O/r02_mode_costs.jsonl 2> $O/r02_mode_costs.err
tail -2 $O/r02_mode_costs.jsonl
