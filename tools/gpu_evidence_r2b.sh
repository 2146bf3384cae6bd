#!/bin/bash
# ncu captures of the shipped build for the configs whose kernel changed after tools/gpu_evidence_r2.sh: cfg5 (streamed, tensor path) and `random` (moving spheres, tensor path)
out=gpurun_out; mkdir -p $out
cap() {  # cap <tag> <flavour> <spp> <preset> <w> <h> <kernel regex>
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$7 -c 1 -o $out/ncu_r2_$1 -f python tools/wave_one.py $2 $3 $4 $5 $6 > $out/ncu_r2_$1.log 2>&1; echo "$1 exit $?"; tail -1 $out/ncu_r2_$1.log
}
cap cfg5_streamed_mma_4spp 0 4 stress100k 1920 1080 pt_megakernel
cap random_mma_128spp 0 128 random 1200 800 pt_megakernel
