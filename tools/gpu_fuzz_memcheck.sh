#!/bin/bash
# developer tool: tools/gpu_fuzz.py under compute-sanitizer memcheck (device-side out-of-bounds reads do not raise CUDA errors on their own)
mkdir -p gpurun_out
for sc in cornell_box random_scene cornell_box_smoke oneweek_final textured_lit; do
  timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/gpu_fuzz.py $sc ${SEED:-3} ${TRIALS:-60} > gpurun_out/fuzz_memcheck_${sc}.out 2>&1
  echo "$sc rc=$? $(grep -c 'Invalid' gpurun_out/fuzz_memcheck_${sc}.out) invalid; $(tail -2 gpurun_out/fuzz_memcheck_${sc}.out | tr '\n' ' ' | cut -c1-250)"
done
