#!/bin/bash
# cfg2 at full spp for several chunk sizes (PTGPU_CHUNK_SAMPLES).  usage: tools/gpu_chunks.sh <tag> [variant]
tag=$1; v=$2
out=gpurun_out; mkdir -p $out
for c in 16 32 64 128 256 0; do
  PTGPU_CHUNK_SAMPLES=$c timeout 120 python tools/variant_bench.py "$v" 1024 random_spheres 1200 800 3 2>&1 | tail -1 | tee -a $out/chunks_$tag.txt
done
