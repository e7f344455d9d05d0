#!/bin/bash
# Build avt_b200/variants/libavt_<name>.so = the current objects with ONE source recompiled under extra flags (A/B experiments:
# AVT_B200_LIB=avt_b200/variants/libavt_<name>.so python tools/time_attn.py). Usage: tools/build_variant.sh name file.cu -DFLAG ...
set -e
name=$1; src=$2; shift 2
mkdir -p avt_b200/variants avt_b200/build/variants
obj=avt_b200/build/variants/${name}.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c avt_b200/csrc/$src -o $obj
others=$(ls avt_b200/build/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o avt_b200/variants/libavt_${name}.so $obj $others -gencode arch=compute_100a,code=sm_100a
echo avt_b200/variants/libavt_${name}.so
