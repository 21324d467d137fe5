#!/bin/sh
# Builds instrumented copies of the library (perf probes, wrong results by design) next to the real one:
#   pytorch_probabilisticlayers_b200/lib/probe_<name>/libpl_b200.so     use with PL_B200_LIB=<path>
# usage: tools/build_probe.sh <name> "<extra nvcc -D flags>"
set -e
cd "$(dirname "$0")/../pytorch_probabilisticlayers_b200/csrc"
name=$1; shift
out=../lib/probe_$name; mkdir -p $out ../build/probe_$name
for f in api adam bn_pool ce comm kl linear linear_simt linear_tc conv_simt conv_tc; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    --expt-relaxed-constexpr "$@" -c $f.cu -o ../build/probe_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpl_b200.so ../build/probe_$name/*.o -lcudart -lcuda
echo built $out/libpl_b200.so
