#!/bin/bash
# streamed kernel with the tensor-path pre-filter: tests that touch the streamed kernel + cfg5 A/B (8 spp, full resolution)
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_hits.py tests/test_gpu_parity.py -x -q -m gpu -k "stream or cfg5 or stress or flavour or tensor" > $out/pytest_gpu_r2_mma_streamed.log 2>&1; echo "pytest exit $?"; tail -6 $out/pytest_gpu_r2_mma_streamed.log
: > $out/flavours_r2_cfg5.txt
for rk in 4 5; do timeout 300 python tools/flavour_bench.py '' $rk 8 stress100k 1920 1080 >> $out/flavours_r2_cfg5.txt 2>&1; done
cat $out/flavours_r2_cfg5.txt
