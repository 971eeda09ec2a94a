#!/bin/bash
# Round-2 ncu captures (one GPU).  Output: gpurun_out/r2_*.  Summaries are made from them by tools/summarise_profiles.py.
set -x
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-general --no-blocks --no-check --no-full-d2h"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_bench_under_ncu.log 2>&1
S="python tools/sweep_fill.py --iters 2 --only default"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fill_brick -s 4 -c 1 -o $O/r2_prof_brick $S > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fill_rowtile -s 4 -c 1 -o $O/r2_prof_boundary $S > /dev/null 2>&1
G="python tools/sweep_fill.py --iters 2 --only default --perturb 0.2"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_elem_general -s 2 -c 1 -o $O/r2_prof_elem_general $G > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fill_rowtile -s 2 -c 1 -o $O/r2_prof_rows_general $G > /dev/null 2>&1
K="python tools/bench_blocks.py --steps 1 --n-elastic 96 --n-hcurl 96"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gblock_q2_dmma -s 1 -c 1 -o $O/r2_prof_q2_dmma $K > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_gblock<1, 27" -s 1 -c 1 -o $O/r2_prof_q2_dfma $K > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_gblock<2, 8" -s 1 -c 1 -o $O/r2_prof_elasticity $K > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_gblock<3, 12" -s 1 -c 1 -o $O/r2_prof_hcurl $K > /dev/null 2>&1
ls -la $O | tail -20
