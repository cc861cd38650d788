#!/bin/bash
# Capture `ncu --set full` for a list of kernels on one bench.py workload and keep only compact summaries
# (gpurun_out/ is limited to 64 MiB): <out>/<kernel>.raw.csv (all metrics, one launch), <out>/<kernel>.src.txt (the SASS
# instructions with most warp-stall samples) and the gzipped SASS-level source page.
# usage: tools/ncu_capture.sh <workload> <outdir> <skip> kernel-regex1 kernel-regex2 ...     (env NCU_EXTRA: more ncu flags)
set -u
WL=$1; OUT=$2; SKIP=$3; shift 3
mkdir -p "$OUT"
for k in "$@"; do
  tag=$(echo "$k" | tr -c 'A-Za-z0-9_\n' '_')
  rep=/tmp/prof_$tag
  timeout 300 ncu --set full --clock-control none ${NCU_EXTRA:-} --import-source on -k "regex:$k" -s "$SKIP" -c 1 -f -o "$rep" \
      python bench.py --workload "$WL" --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
  if [ -f "$rep.ncu-rep" ]; then
    ncu -i "$rep.ncu-rep" --page raw --csv > "$OUT/$tag.raw.csv" 2>/dev/null
    ncu -i "$rep.ncu-rep" --page source --csv > "/tmp/src_$tag.csv" 2>/dev/null
    python tools/ncu_src_top.py "/tmp/src_$tag.csv" 40 > "$OUT/$tag.src.txt" 2>&1
    gzip -c "/tmp/src_$tag.csv" > "$OUT/$tag.src.csv.gz"
    rm -f "$rep.ncu-rep" "/tmp/src_$tag.csv"
  else
    echo "no report for $k" > "$OUT/$tag.raw.csv"
  fi
done
ls -la "$OUT"
