#!/bin/bash
# One gpurun call: ncu evidence of the SHIPPED build (default kernels) for every BASELINE config + launch list.  Outputs under gpurun_out/.
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_r2.txt 2>&1
cap() {  # cap <tag> <flavour> <spp> <preset> <w> <h> <kernel regex>
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$7 -c 1 -o $out/ncu_r2_$1 -f python tools/wave_one.py $2 $3 $4 $5 $6 > $out/ncu_r2_$1.log 2>&1; echo "$1 exit $?"; tail -1 $out/ncu_r2_$1.log
}
cap cfg2_regroup 0 1024 random_spheres 1200 800 pt_megakernel
cap cfg4_regroup_64spp 0 64 random_spheres 3840 2160 pt_megakernel
cap cfg3_regroup_256spp 0 256 two_perlin_spheres 1920 1080 pt_megakernel
cap cfg5_streamed_4spp 0 4 stress100k 1920 1080 pt_megakernel
cap random_regroup_128spp 0 128 random 1200 800 pt_megakernel
echo "== ncu launch list of the bench command"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r2.csv python bench.py --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_launch_r2.log 2>&1; echo "exit $?"
ls -la $out/*.ncu-rep
