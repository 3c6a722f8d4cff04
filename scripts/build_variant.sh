#!/bin/bash
# Development aid: build libmcl_b200 with extra -D flags into build/variants/lib<name>.so (git-ignored; travels to the
# GPU box).  Usage: scripts/build_variant.sh <name> [-DFLAG=V ...];  run with MCL_B200_LIB=build/variants/lib<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared \
  -ccbin /usr/bin/g++ "$@" -Xptxas -v -o build/variants/lib$name.so \
  monte_carlo_localization_b200/csrc/mcl_b200.cu monte_carlo_localization_b200/csrc/map_prep.cpp 2> build/variants/$name.ptxas.log
grep -A2 "k_raycast_dirILi207" build/variants/$name.ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo
