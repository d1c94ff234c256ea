#!/bin/bash
# Wall time of the device-resident detection step (4096 frames x 200 candidates, batches overlapped as in bench.py) for the default
# library and every A/B build under build/ (tools/build_variants.sh).  Run on the GPU box.
cd "$(dirname "$0")/.."
echo "== default"; python tools/prof_step.py --frames 4096 --steps 10 --wall | head -1
for lib in build/libtsd_*.so; do
  echo "== $lib"; TSD_LIB=$PWD/$lib python tools/prof_step.py --frames 4096 --steps 10 --wall | head -1
done
