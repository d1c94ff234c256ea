#!/bin/bash
# A/B of the fold on the GPU box: dedup / chain parity tests with the default library, then stage times (detection 4096 frames,
# recognition 1024 frames, 4K x 2000 candidates) and overlapped wall times for the default library and every build/libtsd_*.so.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; L=gpurun_out/${NAME:-abfold}.log; : > $L
[ -z "$SKIP_TESTS" ] && (python -m pytest tests -m gpu -x -q -k "dedup or chain or fuzz or kat or recognition or slot or f64" 2>&1 | tail -2) >> $L
for lib in "" build/libtsd_*.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== ${lib:-default}" >> $L
  for args in "--frames 4096" "--mode rec --frames 1024" "--H 2160 --W 3840 --boxes 2000 --frames 256"; do
    TSD_LIB=${lib:+$PWD/$lib} python tools/prof_step.py $args --steps 5 --times 2>&1 | tail -2 | head -1 | sed "s/.*'k5_fold': \([0-9.]*\).*/fold \1/" | tr '\n' ' ' >> $L
    TSD_LIB=${lib:+$PWD/$lib} python tools/prof_step.py $args --steps 10 --wall 2>&1 | head -1 >> $L
  done
done
cat $L
