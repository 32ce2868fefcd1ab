#!/bin/bash
# usage: tools/build_variant.sh NAME "-DRBP_OPF=1 ..."   -> analog-pde-solver-sim_b200/lib/libapde_NAME.so (tuning aid)
# Rebuilds the four red-black kernel translation units with extra flags; run with APDE_LIBRARY=<that .so>.
set -e
cd "$(dirname "$0")/../analog-pde-solver-sim_b200/csrc"
NAME=$1; FLAGS=$2
COMMON="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../include -I. -Xcompiler -fPIC -Xptxas -v"
mkdir -p build/var_$NAME
for f in apde_rb_stream_f64 apde_rb_stream_f32 apde_rb_pipe_f64 apde_rb_pipe_f32; do
  nvcc $COMMON $FLAGS -c $f.cu -o build/var_$NAME/$f.o 2> build/var_$NAME/$f.ptxas.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libapde_$NAME.so build/apde_exact.o build/apde_redblack.o \
  build/var_$NAME/apde_rb_stream_f64.o build/var_$NAME/apde_rb_stream_f32.o build/var_$NAME/apde_rb_pipe_f64.o build/var_$NAME/apde_rb_pipe_f32.o \
  build/apde_multigrid.o build/apde_api.o -Xcompiler -fPIC
echo built $NAME
