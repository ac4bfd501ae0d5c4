#!/bin/bash
# Build kernel variants of libaudiomps.so for A/B timing: profiles/build_variants.sh name "-DFLAG=..." ...
# (the product build is audio_mps_b200/_lib.py:build(); select a variant at run time with AMPS_LIB=path)
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o profiles/variants/lib_$name.so audio_mps_b200/csrc/amps_api.cu
echo built profiles/variants/lib_$name.so
