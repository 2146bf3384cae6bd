#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over small renders of every kernel flavour (development aid).
tag=$1; out=gpurun_out; mkdir -p $out
python - <<'PY'
import numpy as np, pathtrace_rs_b200 as pt
pt.write_ppm("/tmp/earth.ppm", (np.arange(32*16*3) % 251).astype(np.uint8).reshape(16, 32, 3))
PY
export PATHTRACE_EARTHMAP=/tmp/earth.ppm
run() { # tool, label, env..., -- preset w h spp
  tool=$1; shift; label=$1; shift
  echo "== $tool $label" | tee -a $out/sanitize_$tag.txt
  env "$@" timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 python tools/variant_bench.py "" $SPP $PRESET $W $H 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|variant|Error|hazard" | head -8 | tee -a $out/sanitize_$tag.txt
}
for tool in memcheck racecheck synccheck; do
  PRESET=random_spheres W=64 H=32 SPP=4 run $tool resident X=1
  PRESET=random W=64 H=32 SPP=4 run $tool resident_motion X=1
  PRESET=two_perlin_spheres W=64 H=32 SPP=4 run $tool noise X=1
  PRESET=earth W=64 H=32 SPP=4 run $tool image X=1
  PRESET=random_spheres W=64 H=32 SPP=2 run $tool streamed PTGPU_FORCE_STREAM_TILE_BLOCKS=16
  PRESET=random_spheres W=200 H=120 SPP=16 run $tool chunked PTGPU_CHUNK_SAMPLES=4
done
