#!/bin/bash
# developer tool: build tuning variants of libtpt.so into gpurun-visible dirs
cd "$(dirname "$0")/../tiny-path-tracer_b200/csrc"
i=0
while read -r flags; do
  d="$(cd ../.. && pwd)/gpurun_variants/v$i"
  mkdir -p $d
  make -s -j3 LIB=$d EXTRA="$flags" >/dev/null 2>&1 || echo "build failed: $flags"
  echo "$flags" > $d/flags.txt
  grep -E "render_wave_kernelILb0ELb1ELb1ELb0ELb0ELb0ELb1" -A2 $d/obj/ptxas_fast.log | grep -E "registers" | head -1
  i=$((i+1))
done
