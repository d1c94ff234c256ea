#!/bin/bash
# Grid-size A/B on the GPU box: per-stage times (detection, 4096 frames) for build/libtsd_*.so and for TSD_K2_GRID values.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; L=gpurun_out/${NAME:-abgrid}.log; : > $L
for lib in "" build/libtsd_*.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== ${lib:-default}" >> $L
  TSD_LIB=${lib:+$PWD/$lib} python tools/prof_step.py --frames 4096 --steps 5 --times 2>&1 | tail -2 | head -1 >> $L
done
for g in 12 24 48 96; do
  echo "== TSD_K2_GRID=$g" >> $L
  TSD_K2_GRID=$g python tools/prof_step.py --frames 4096 --steps 5 --times 2>&1 | tail -2 | head -1 >> $L
done
cat $L
