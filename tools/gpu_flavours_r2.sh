#!/bin/bash
# A/B of the resident-kernel flavours on the same box (cfg2 at 256 spp and cfg4 geometry at 16 spp)
out=gpurun_out/flavours_r2.txt
: > $out
for rk in 4 5; do
  timeout 120 python tools/flavour_bench.py '' $rk 256 random_spheres 1200 800 >> $out 2>&1
done
for rk in 4 5; do
  timeout 120 python tools/flavour_bench.py '' $rk 32 random_spheres 3840 2160 >> $out 2>&1
done
for rk in 4 5; do
  timeout 120 python tools/flavour_bench.py '' $rk 128 random 1200 800 >> $out 2>&1
done
cat $out
