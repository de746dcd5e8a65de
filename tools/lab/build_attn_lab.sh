#!/bin/bash
# builds the attention lab variants into tools/lab/bin (git-ignored; travels to the GPU box)
set -e
cd "$(dirname "$0")"
mkdir -p bin
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lcuda -o bin/attn_$1 attn_lab.cu "${@:2}" & }
build base
build poly0 -DS3OD_ATTN_POLY_EVERY=0
build poly2 -DS3OD_ATTN_POLY_EVERY=2
build poly3 -DS3OD_ATTN_POLY_EVERY=3
build nomax -DS3OD_ATTN_LAB=1
build noexp -DS3OD_ATTN_LAB=2
build nomax_noexp -DS3OD_ATTN_LAB=3
build pp -DS3OD_ATTN_PINGPONG=1
build pp_poly0 -DS3OD_ATTN_PINGPONG=1 -DS3OD_ATTN_POLY_EVERY=0
build mmaonly -DS3OD_ATTN_LAB=16
wait
ls bin
