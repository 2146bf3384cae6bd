#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>  ->  pathtrace_rs_b200/lib/<name>/{libptgpu.so,libpthost.so}  (development aid)
set -e
cd "$(dirname "$0")/../pathtrace_rs_b200"
name=$1; shift
mkdir -p lib/$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -shared -o lib/$name/libptgpu.so csrc/ptgpu.cu 2>&1 | grep -E "error|Compiling entry|Used" | grep -B1 "Used" | grep -v "^--" | paste - - | sed -E 's/.*function .(_ZN2pt[0-9]*[a-z_]*).*Used ([0-9]+) registers.*/\1 regs=\2/' 
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -shared -o lib/$name/libpthost.so host/pathtrace.cpp -Llib/$name -lptgpu -Wl,-rpath,'$ORIGIN'
