#!/bin/bash
# developer tool: tools/gpu_diff_fuzz.py over the scene classes (parity mode vs the plain-C restatement on perturbed scenes)
mkdir -p gpurun_out
for sc in cornell_box sphere_cornell_box random_scene random_scene_list light_spheres cornell_box_smoke oneweek_final; do
  timeout 200 python tools/gpu_diff_fuzz.py $sc ${SEED:-1} ${TRIALS:-40} > gpurun_out/diff_fuzz_${sc}.out 2>&1
  echo "$sc rc=$? $(tail -1 gpurun_out/diff_fuzz_${sc}.out | cut -c1-300)"; grep -c "HITS\|RENDER" gpurun_out/diff_fuzz_${sc}_${SEED:-1}.log
done
