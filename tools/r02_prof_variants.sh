#!/bin/bash
for v in "$@"; do
  export TPT_LIBTPT=gpurun_variants/v$v/libtpt.so
  SCENES=random_scene QUICK=1 KERN=1 bash tools/gpu_prof_scenes.sh
  python profiles/tools/ncu_lines.py gpurun_out/scene_random_scene_k1.ncu-rep 70 > gpurun_out/r02_vote_v${v}_lines.txt 2>&1
  python profiles/tools/ncu_scene_summary.py v$v=gpurun_out/scene_random_scene_k1.ncu-rep > gpurun_out/r02_vote_v${v}_summary.json
  rm -f gpurun_out/scene_random_scene_k1.ncu-rep
done
