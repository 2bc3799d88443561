#!/bin/bash
# Compile one translation unit with -Xptxas -v and print registers/spills per kernel.
# Usage: tools/ptxas_info.sh snowtri_p1.cu [filter]
src=snowmocap_b200/csrc/$1
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -I include -I snowmocap_b200/csrc -Xcompiler -fPIC -Xptxas=-v -c $src -o /tmp/ptxas_info.o 2>&1 | \
  awk '/Compiling entry function/ {name=$5} /bytes stack frame/ {spill=$0} /Used [0-9]+ registers/ {print name, $0, spill}' | \
  sed -e 's/ptxas info    ://g' -e "s/for 'sm_100a'//" | c++filt | grep -E "${2:-.}" | cut -c1-260
