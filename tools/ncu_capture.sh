#!/bin/bash
# Capture `ncu --set full` for a list of kernels on one bench.py workload and keep only compact text
# summaries (gpurun_out/ is limited to 64 MiB): <out>/<kernel>.raw.csv (all metrics, one launch) and
# <out>/<kernel>.src.txt (the 40 source lines with most warp-stall samples).
# usage: tools/ncu_capture.sh <workload> <outdir> <skip> kernel1 kernel2 ...
set -u
WL=$1; OUT=$2; SKIP=$3; shift 3
mkdir -p "$OUT"
for k in "$@"; do
  rep=/tmp/prof_$k
  timeout 300 ncu --set full --clock-control none ${NCU_EXTRA:-} --import-source on -k regex:$k -s "$SKIP" -c 1 -f -o "$rep" \
      python bench.py --workload "$WL" --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
  if [ -f "$rep.ncu-rep" ]; then
    ncu -i "$rep.ncu-rep" --page raw --csv > "$OUT/$k.raw.csv" 2>/dev/null
    ncu -i "$rep.ncu-rep" --page source --csv > /tmp/src_$k.csv 2>/dev/null
    python tools/ncu_src_top.py /tmp/src_$k.csv 40 > "$OUT/$k.src.txt" 2>&1
    gzip -c /tmp/src_$k.csv > "$OUT/$k.src.csv.gz"
    rm -f "$rep.ncu-rep" /tmp/src_$k.csv
  else
    echo "no report for $k" > "$OUT/$k.raw.csv"
  fi
done
ls -la "$OUT"
