#!/bin/bash
# Overlap experiments on the GPU box: wall time per device-resident step with the staggered two-slot pipeline on / off and the cap of
# resident fold CTAs per SM.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; L=gpurun_out/${NAME:-foldcap}.log; : > $L
(python -m pytest tests -m gpu -x -q -k "back_to_back or chain_synthetic or fetch_previous or overlap" 2>&1 | tail -2) >> $L
for st in 0 1; do for cap in 0 1 2 3; do
  echo "== TSD_STAGGER=$st TSD_FOLD_PER_SM=$cap: det 4096 / det 1024 / rec 1024" >> $L
  TSD_STAGGER=$st TSD_FOLD_PER_SM=$cap python tools/prof_step.py --frames 4096 --steps 12 --wall 2>&1 | head -1 >> $L
  TSD_STAGGER=$st TSD_FOLD_PER_SM=$cap python tools/prof_step.py --frames 1024 --steps 20 --wall 2>&1 | head -1 >> $L
  TSD_STAGGER=$st TSD_FOLD_PER_SM=$cap python tools/prof_step.py --mode rec --frames 1024 --steps 12 --wall 2>&1 | head -1 >> $L
done; done
cat $L
