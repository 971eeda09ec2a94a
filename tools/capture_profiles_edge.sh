#!/bin/bash
# Round-2 (late) ncu captures after k_fill_edge became the boundary kernel (one GPU).  Output: gpurun_out/r2b_*.
set -x
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-general --no-blocks --no-check --no-full-d2h"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches_bench.csv $B > $O/r2b_bench_under_ncu.log 2>&1
S="python tools/sweep_fill.py --iters 2 --only default"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fill_edge -s 4 -c 1 -o $O/r2b_prof_edge $S > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fill_brick -s 4 -c 1 -o $O/r2b_prof_brick $S > /dev/null 2>&1
ls -la $O | tail -8
