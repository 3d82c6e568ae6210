#!/bin/bash
# tools/build_variant.sh NAME "<extra nvcc flags>": a development variant of libb200icp.so
# -> mola-fe-lidar_b200/lib/var_NAME.so (select with B200ICP_LIB=<path>)
set -e
NAME=$1; shift
EXTRA="$*"
cd "$(dirname "$0")/../mola-fe-lidar_b200"
B=build/var/$NAME; mkdir -p $B lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall --expt-relaxed-constexpr -I../include $EXTRA"
for f in runtime cloud align voxel edges_planes capi; do
  /usr/local/cuda/bin/nvcc $FLAGS -Xptxas -v -c csrc/$f.cu -o $B/$f.o 2> $B/$f.log &
done
/usr/local/cuda/bin/nvcc $FLAGS -x cu -c csrc/icp_params_yaml.cpp -o $B/icp_params_yaml.o 2> $B/yaml.log &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/var_$NAME.so $B/*.o build/yaml_lite.o -cudart static
echo "built lib/var_$NAME.so"; grep -A2 "search_tile_kernelILi6ENS_8NnWriter" $B/align.log | grep -E "Used|spill"
