#!/bin/bash
# usage: tools/build_variant.sh NAME "-DRBS_PF_ROWS=4 ..."   -> analog-pde-solver-sim_b200/lib/libapde_NAME.so (tuning aid)
set -e
cd "$(dirname "$0")/../analog-pde-solver-sim_b200/csrc"
NAME=$1; FLAGS=$2
COMMON="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../include -I. -Xcompiler -fPIC"
mkdir -p build/var_$NAME
nvcc $COMMON $FLAGS -c apde_rb_stream_f64.cu -o build/var_$NAME/f64.o &
nvcc $COMMON $FLAGS -c apde_rb_stream_f32.cu -o build/var_$NAME/f32.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libapde_$NAME.so build/apde_exact.o build/apde_redblack.o build/var_$NAME/f64.o build/var_$NAME/f32.o build/apde_api.o -Xcompiler -fPIC
echo built $NAME
