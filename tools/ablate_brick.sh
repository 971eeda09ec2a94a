#!/bin/bash
# Ablation of k_fill_brick on the 256^3 workload: TXASM_BRICK_ABLATE bit mask (1 no gathers, 2 no stencil phase, 4 no stores);
# results are wrong by construction, only the time matters.  Output: gpurun_out/r2_ablate_brick.txt
out=gpurun_out/r2_ablate_brick.txt
: > $out
for a in ${ABLATE_SET:-0 1 2 3 4 5 6 7}; do
  echo "== TXASM_BRICK_ABLATE=$a" >> $out
  TXASM_BRICK_ABLATE=$a timeout 200 python tools/sweep_fill.py --iters 20 --only "default,no concurrent,brick 3 CTA/SM,brick 2 CTA/SM" 2>&1 | grep -v "^setup" >> $out
done
