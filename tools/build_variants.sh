#!/bin/bash
# A/B builds of the library with build-time knobs: tools/build_variants.sh name "-DX=1 -DY=2" ...  -> build/libtsd_<name>.so
# (run a script with TSD_LIB=build/libtsd_<name>.so to use one).  build/ is git-ignored and travels to the GPU box.
set -e
cd "$(dirname "$0")/.."
mkdir -p build
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -s -B -C opencv-traffic-sign-detector_b200/csrc OUT=../../build/libtsd_$name.so EXTRA="$flags" &
done
wait
ls -la build/
