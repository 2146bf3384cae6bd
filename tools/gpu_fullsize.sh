#!/bin/bash
# Full-size bench lines of the other BASELINE configs (cfg2 is the default bench line).  usage: tools/gpu_fullsize.sh <tag>
tag=$1
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q -k "cfg3_full or image or earth" 2>&1 | tail -2
timeout 300 python bench.py --workload cfg1 > $out/bench_${tag}_cfg1.json 2>$out/bench_${tag}_cfg1.err; tail -c 300 $out/bench_${tag}_cfg1.json; echo
timeout 300 python bench.py --workload cfg3 > $out/bench_${tag}_cfg3.json 2>$out/bench_${tag}_cfg3.err; tail -c 300 $out/bench_${tag}_cfg3.json; echo
timeout 600 python bench.py --workload cfg4 --fast --no-cpu-baseline > $out/bench_${tag}_cfg4_n1.json 2>$out/bench_${tag}_cfg4.err; tail -c 300 $out/bench_${tag}_cfg4_n1.json; echo
timeout 900 python bench.py --workload cfg5 --fast --no-cpu-baseline > $out/bench_${tag}_cfg5.json 2>$out/bench_${tag}_cfg5.err; tail -c 300 $out/bench_${tag}_cfg5.json; echo
