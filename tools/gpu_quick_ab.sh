#!/bin/bash
# quick A/B of the current build: hit tests + the three flavour_bench configurations + cfg5 at 8 spp
out=gpurun_out/quick_ab.txt; : > $out
timeout 600 python -m pytest tests/test_gpu_hits.py -x -q -m gpu 2>&1 | tail -2 >> $out
timeout 120 python tools/flavour_bench.py '' 5 32 random_spheres 3840 2160 >> $out 2>&1
timeout 120 python tools/flavour_bench.py '' 5 256 random_spheres 1200 800 >> $out 2>&1
timeout 120 python tools/flavour_bench.py '' 5 128 random 1200 800 >> $out 2>&1
timeout 300 python tools/flavour_bench.py '' 5 8 stress100k 1920 1080 >> $out 2>&1
cat $out
