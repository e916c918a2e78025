#!/bin/bash
# developer tool: BVH-scene throughput of every tuning build under gpurun_variants/
for d in gpurun_variants/v*; do
  echo "== $(cat $d/flags.txt)"
  TPT_LIBTPT=$d/libtpt.so python tools/gpu_bvh_perf.py ${1:-32} 2>&1 | grep "kernel ${KERN:-[01]}" | grep -v smoke
done
