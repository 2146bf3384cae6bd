#!/bin/bash
# A/B of build variants of the tensor-path regroup kernel (same box, same call): python tools/flavour_bench.py <variant> <flavour> ...
out=gpurun_out/variants_r2_mma.txt
: > $out
for v in "" p2 p3 cta128 cta128p2; do
  timeout 120 python tools/flavour_bench.py "$v" 5 32 random_spheres 3840 2160 >> $out 2>&1
  timeout 120 python tools/flavour_bench.py "$v" 5 256 random_spheres 1200 800 >> $out 2>&1
  timeout 120 python tools/flavour_bench.py "$v" 5 128 random 1200 800 >> $out 2>&1
done
cat $out
