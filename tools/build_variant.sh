#!/bin/bash
# tools/build_variant.sh NAME -DFLAG...   ->  petastorm_b200/variants/libpst_NAME.so  (kernels_decode.cu rebuilt with the flags)
set -e
cd "$(dirname "$0")/../petastorm_b200/csrc"
name=$1; shift
mkdir -p ../variants build_v
NVF="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function -cudart shared"
/usr/local/cuda/bin/nvcc $NVF "$@" -c kernels_decode.cu -o build_v/kernels_decode_$name.o
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart shared -o ../variants/libpst_$name.so build/parquet_meta.o build/host_api.o build_v/kernels_decode_$name.o build/kernels_copy.o build/kernels_ops.o build/kernels_png.o build/ctx.o build/jpeg.o -Xlinker -rpath,/usr/local/cuda/lib64 -ldl -lpthread
