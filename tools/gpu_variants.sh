#!/bin/bash
# Times kernel variants (tools/build_variant.sh builds) on cfg2 / cfg3 / random at reduced spp.  usage: tools/gpu_variants.sh <tag> <variants...>
tag=$1; shift
out=gpurun_out; mkdir -p $out
for v in "" "$@"; do
  timeout 120 python tools/variant_bench.py "$v" 256 random_spheres 1200 800 3 2>&1 | tail -1 | tee -a $out/variants_$tag.txt
  timeout 120 python tools/variant_bench.py "$v" 64 two_perlin_spheres 1920 1080 3 2>&1 | tail -1 | tee -a $out/variants_$tag.txt
  timeout 120 python tools/variant_bench.py "$v" 128 random 1200 800 3 2>&1 | tail -1 | tee -a $out/variants_$tag.txt
done
