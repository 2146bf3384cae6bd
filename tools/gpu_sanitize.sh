#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over small renders of every kernel flavour.  usage: tools/gpu_sanitize.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out; : > $out/sanitize_$tag.txt
for tool in memcheck synccheck racecheck; do
  for c in ${CASES:-mma mma_motion mma_chunked regroup regroup_motion noise image streamed streamed_motion chunked pair_const pair_lds wave wave_motion_chunked multi debug_hits}; do
    # racecheck understands barriers, not the wavefront kernel's lock-free queues (flag-guarded hand-offs it reports as hazards)
    if [ $tool = racecheck ] && [[ $c == wave* ]]; then continue; fi
    echo "== $tool $c" | tee -a $out/sanitize_$tag.txt
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_case.py $c 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case |Error|hazard|Traceback" | head -6 | tee -a $out/sanitize_$tag.txt
  done
done
