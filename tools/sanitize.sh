#!/bin/bash
# compute-sanitizer on a small device-resident chain + the host path (one tool per gpurun call: B200_PROFILING.md).
# usage: tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
set -e
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
python tools/prof_step.py --frames 24 --steps 1 > gpurun_out/san_plain.log 2>&1
/usr/local/cuda/bin/compute-sanitizer --tool "$tool" --error-exitcode 3 python tools/prof_step.py --frames 24 --steps 1 > gpurun_out/san_${tool}.log 2>&1
tail -5 gpurun_out/san_${tool}.log
