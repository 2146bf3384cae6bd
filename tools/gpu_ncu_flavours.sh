#!/bin/bash
# ncu --set full of the two-path resident flavours at cfg4's geometry (32 spp), for comparison with the regroup capture
out=gpurun_out; mkdir -p $out
cap() {  # cap <tag> <flavour> <spp> <preset> <w> <h> <kernel regex>
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$7 -c 1 -o $out/ncu_r2_$1 -f python tools/wave_one.py $2 $3 $4 $5 $6 > $out/ncu_r2_$1.log 2>&1; echo "$1 exit $?"; tail -1 $out/ncu_r2_$1.log
}
cap cfg4_twopath_lds_32spp 3 32 random_spheres 3840 2160 pt_megakernel
cap cfg4_twopath_ur_32spp 2 32 random_spheres 3840 2160 pt_megakernel
