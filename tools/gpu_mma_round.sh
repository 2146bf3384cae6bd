#!/bin/bash
# full GPU test suite on the tensor-path default + an ncu capture of the new kernel at cfg4's geometry (64 spp)
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest_gpu_r2_mma.log 2>&1; echo "pytest exit $?"; tail -5 $out/pytest_gpu_r2_mma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pt_megakernel -c 1 -o $out/ncu_r2_cfg4_mma_64spp -f python tools/wave_one.py 0 64 random_spheres 3840 2160 > $out/ncu_r2_cfg4_mma_64spp.log 2>&1; echo "ncu exit $?"; tail -1 $out/ncu_r2_cfg4_mma_64spp.log
