#!/bin/bash
# developer tool (runs HERE, after tools/gpu_round.sh came back): turn gpurun_out/ into the committed summaries under profiles/
set -e
cd "$(dirname "$0")/.."
O=gpurun_out; P=profiles
for f in $O/r02_bench_config*_n*.json $O/r02_bench_ref_n1.json $O/r02_pytest_gpu.txt $O/r02_smoke.txt $O/r02_launches_bench_steps2.csv \
         $O/r02_sanitizer_memcheck.log $O/r02_sanitizer_racecheck.log $O/r02_fast_mismatch.json $O/r02_pytest_multi_n*.txt; do
  [ -f "$f" ] && cp "$f" $P/ || true
done
cp $O/r02_ncu_full_*_raw.csv $O/r02_ncu_full_*_details.txt $O/r02_ncu_full_*_lines.txt $O/r02_scene_ncu_summary.json $P/
# bench.py's roofline reads profiles/ncu_summary.json: captured from bench.py --spp 256 (fast, config 4A)
python $P/tools/ncu_summary.py $P/r02_ncu_full_render_wave_fast_A_raw.csv 368640000 "ncu --set full --clock-control none, bench.py --spp 256 --steps 1 --warmup 1 (launch 2 of render_wave_kernel); the framebuffer traffic of a launch does not depend on spp: R x npix x 12 B" 4A fast > $P/ncu_summary.json
python $P/tools/ncu_summary.py $P/r02_ncu_full_render_wave_parity_A_raw.csv 92160000 "ncu --set full, bench.py --mode parity --spp 64" 4A parity > $P/r02_ncu_summary_parity.json
ls -la $P | grep r02 | wc -l
