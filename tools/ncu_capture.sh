#!/bin/bash
# One `ncu --set full` capture on the GPU box, reduced there to text (gpurun_out/ is capped at 64 MiB, a report with sources is 25-40 MB):
#   tools/ncu_capture.sh <name> "<kernel regex>" <max launches> "<kernels to break down by source line, space separated>" -- <command...>
# writes gpurun_out/<name>.summary.txt (tools/ncu_summary.py: one line per launch) and gpurun_out/<name>.<kernel>.lines.txt
# (tools/ncu_lines.py: stall samples and executed warp instructions per CUDA source line); the report itself is deleted unless KEEP_REP=1.
name=$1; regex=$2; count=$3; lines=$4; shift 5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$regex" -c "$count" -o gpurun_out/$name -f "$@" > gpurun_out/$name.ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.summary.txt 2>&1
for k in $lines; do
  ncu -i gpurun_out/$name.ncu-rep --page source --print-source cuda,sass --csv -k regex:"$k" > gpurun_out/$name.$k.source.csv 2>/dev/null
  python tools/ncu_lines.py gpurun_out/$name.$k.source.csv 45 > gpurun_out/$name.$k.lines.txt 2>&1
  rm -f gpurun_out/$name.$k.source.csv
done
[ -n "$TRAFFIC_WINDOWS" ] && python tools/make_traffic.py gpurun_out/$name.ncu-rep "$TRAFFIC_WINDOWS" "ncu --set full, $*, profiles/${TRAFFIC_LABEL:-$name}" gpurun_out/$name.traffic.json > /dev/null 2>&1
[ "$KEEP_REP" = "1" ] || rm -f gpurun_out/$name.ncu-rep
tail -3 gpurun_out/$name.ncu.log
