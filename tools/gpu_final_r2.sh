#!/bin/bash
# final rehearsal of what the driver runs at round end on one GPU: smoke(), the bench at its default size, the reference arm
out=gpurun_out; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r2_final.txt 2>&1; echo "smoke exit $?"; tail -2 $out/smoke_r2_final.txt
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_r2_v6_cfg4_n1.json 2> $out/bench_r2_v6_n1.err; echo "bench exit $? lines $(wc -l < $out/bench_r2_v6_cfg4_n1.json)"; cut -c1-260 $out/bench_r2_v6_cfg4_n1.json
