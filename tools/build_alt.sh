#!/bin/bash
# build a variant of the library with extra -D flags into build/alt/<name>.so (perf experiments on the GPU box swap it in)
set -e
name=$1; shift
mkdir -p build/alt
NVF="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -diag-suppress 177,550"
nvcc $NVF "$@" -c zultra_b200/csrc/zb_capi.cu -o build/alt/$name.capi.o
nvcc $NVF "$@" -c zultra_b200/csrc/zb_prims.cu -o build/alt/$name.prims.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/alt/$name.so build/alt/$name.capi.o build/alt/$name.prims.o build/libzultra.o build/frame.o build/dictionary.o -cudart static
rm -f build/alt/$name.capi.o build/alt/$name.prims.o
ls -la build/alt/$name.so
