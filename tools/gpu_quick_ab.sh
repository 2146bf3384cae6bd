#!/bin/bash
# quick A/B of build variants (tools/build_variant.sh <name> <flags>): usage tools/gpu_quick_ab.sh "<variant> <variant> ..." ('' = the shipped build)
out=gpurun_out/quick_ab.txt; : > $out
for v in "" $1; do
  timeout 120 python tools/flavour_bench.py "$v" 5 32 random_spheres 3840 2160 >> $out 2>&1
  timeout 120 python tools/flavour_bench.py "$v" 5 256 random_spheres 1200 800 >> $out 2>&1
  timeout 300 python tools/flavour_bench.py "$v" 5 8 stress100k 1920 1080 >> $out 2>&1
done
cat $out
