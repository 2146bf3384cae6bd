#!/bin/bash
# evidence of the FINAL round-2 build (tensor-path regroup kernel with the exact blocks in shared memory): full captures of cfg2
# (full size) and cfg4's geometry (64 spp), and the launch list of the bench command at reduced spp
out=gpurun_out; mkdir -p $out
cap() {  # cap <tag> <flavour> <spp> <preset> <w> <h> <kernel regex>
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$7 -c 1 -o $out/ncu_r2_$1 -f python tools/wave_one.py $2 $3 $4 $5 $6 > $out/ncu_r2_$1.log 2>&1; echo "$1 exit $?"; tail -1 $out/ncu_r2_$1.log
}
cap cfg2_final 0 1024 random_spheres 1200 800 pt_megakernel
cap cfg4_final_64spp 0 64 random_spheres 3840 2160 pt_megakernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r2_final.csv python bench.py --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_launch_r2_final.log 2>&1; echo "launch list exit $?"
