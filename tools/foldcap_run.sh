#!/bin/bash
# Overlap experiments on the GPU box: wall time per device-resident step (two-slot pipeline) for the default library and every
# build/libtsd_*.so, over the cap of resident fold CTAs per SM; plus the fold's own time with the batches serialised.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; L=gpurun_out/${NAME:-foldcap}.log; : > $L
for lib in "" build/libtsd_*.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  for cap in ${CAPS:-2 3 4}; do
    echo "== ${lib:-default} TSD_FOLD_PER_SM=$cap: wall det 4096 / stage times / wall rec 1024" >> $L
    TSD_LIB=${lib:+$PWD/$lib} TSD_FOLD_PER_SM=$cap python tools/prof_step.py --frames 4096 --steps 12 --wall 2>&1 | head -1 >> $L
    TSD_LIB=${lib:+$PWD/$lib} TSD_FOLD_PER_SM=$cap python tools/prof_step.py --frames 4096 --steps 5 --times 2>&1 | tail -2 | head -1 >> $L
    TSD_LIB=${lib:+$PWD/$lib} TSD_FOLD_PER_SM=$cap python tools/prof_step.py --mode rec --frames 1024 --steps 12 --wall 2>&1 | head -1 >> $L
  done
done
cat $L
