#!/bin/bash
# tensor-path checks: new GPU tests + compute-sanitizer on the new kernel's cases
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_hits.py tests/test_gpu_parity.py -x -q -m gpu -k "tensor_path or flavour" > $out/pytest_gpu_r2_mma_new.log 2>&1; echo "pytest exit $?"; tail -8 $out/pytest_gpu_r2_mma_new.log
CASES="mma mma_motion mma_chunked debug_hits" timeout 1500 bash tools/gpu_sanitize.sh r2_mma > /dev/null 2>&1; cat $out/sanitize_r2_mma.txt
