#!/bin/bash
# developer tool: tools/gpu_fuzz.py over every scene class, one bounded process each
mkdir -p gpurun_out
for sc in cornell_box sphere_cornell_box random_scene cornell_box_smoke oneweek_final textured_lit; do
  timeout 150 python tools/gpu_fuzz.py $sc ${SEED:-1} ${TRIALS:-300} > gpurun_out/fuzz_${sc}.out 2>&1
  echo "$sc rc=$? $(tail -1 gpurun_out/fuzz_${sc}.out | cut -c1-300)"; cat gpurun_out/fuzz_${sc}_${SEED:-1}.log | cut -c1-300
done
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
