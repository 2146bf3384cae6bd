#!/bin/bash
# One gpurun call: GPU parity suite, variant timings, bench line, ncu launch list + one full capture.  Outputs under gpurun_out/.
# usage: [PARITY_VARIANTS="v ..."] tools/gpu_round.sh <tag> [variants to time...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_$tag.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest_gpu_$tag.log
for v in "" "$@"; do
  echo "== variant '$v'"
  timeout 120 python tools/variant_bench.py "$v" 256 random_spheres 1200 800 3 2>&1 | tail -1 | tee -a $out/variants_$tag.txt
done
for v in $PARITY_VARIANTS; do
  echo "== parity subset with variant $v"
  PTGPU_LIB_DIR=$PWD/pathtrace_rs_b200/lib/$v timeout 600 python -m pytest tests -m gpu -q -k "bit_exact or cfg1 or random_scenes or golden or moving or image or chunk" > $out/pytest_gpu_${tag}_$v.log 2>&1; echo "exit $?"; tail -2 $out/pytest_gpu_${tag}_$v.log
done
echo "== bench cfg2"; timeout 600 python bench.py > $out/bench_${tag}_cfg2.json 2> $out/bench_${tag}_cfg2.err; tail -c 600 $out/bench_${tag}_cfg2.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_launch_$tag.log 2>&1; echo "exit $?"
echo "== ncu full capture (cfg2, one launch)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_megakernel -c 1 -o $out/ncu_${tag}_cfg2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --fast > $out/ncu_full_$tag.log 2>&1; echo "exit $?"; ls -la $out/*.ncu-rep
