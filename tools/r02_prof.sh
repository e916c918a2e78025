#!/bin/bash
# developer tool: ncu --set full captures (with source) of the headline kernels: fast / parity, Cornell A (and B fast)
mkdir -p gpurun_out
for m in fast parity; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_${m}_A python bench.py --steps 1 --warmup 1 --mode $m --spp ${SPP:-64} --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_${m}_A.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_fast_B python bench.py --steps 1 --warmup 1 --variant B --spp 32 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_fast_B.log 2>&1
ls -la gpurun_out/*.ncu-rep
