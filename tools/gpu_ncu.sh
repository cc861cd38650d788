#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel-regex>:<skip> ...     (Au20 workload, one launch each)
set -u
TAG=$1; shift
for spec in "$@"; do
  k=${spec%%:*}; skip=${spec##*:}
  bash tools/ncu_capture.sh cfg3 gpurun_out/${TAG}_ncu "$skip" "$k" > /dev/null 2>&1
done
ls gpurun_out/${TAG}_ncu
