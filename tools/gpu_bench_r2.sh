#!/bin/bash
# round-2 evidence of the shipped build: the driver's bench command at N=1 (cfg4 full size), the ncu launch list of the same
# command at reduced spp, and a full-size ncu capture of cfg2
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_r2_mma.txt 2>&1
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_r2_v5_cfg4_n1.json 2> $out/bench_r2_v5_n1.err; echo "bench exit $?"; cut -c1-600 $out/bench_r2_v5_cfg4_n1.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $out/bench_r2_v5_ref.json 2>> $out/bench_r2_v5_n1.err; echo "ref exit $?"; cut -c1-400 $out/bench_r2_v5_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r2_mma.csv python bench.py --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_launch_r2_mma.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pt_megakernel -c 1 -o $out/ncu_r2_cfg2_mma -f python tools/wave_one.py 0 1024 random_spheres 1200 800 > $out/ncu_r2_cfg2_mma.log 2>&1; echo "ncu exit $?"; tail -1 $out/ncu_r2_cfg2_mma.log
