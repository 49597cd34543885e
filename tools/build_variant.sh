#!/bin/bash
# builds an experimental variant of the library: tools/build_variant.sh <name> [-DFLAG ...]  ->  lib/libphylocsf_b200_<name>.so
# (selected at run time with PCSF_LIB_VARIANT=<name>; the product library is built by __graft_entry__.build())
cd "$(dirname "$0")/.."
NAME=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static "$@" \
  -o phylocsfpp_b200/lib/libphylocsf_b200_$NAME.so phylocsfpp_b200/csrc/pcsf_capi.cu 2>&1 | grep -i "error" | head
ls -la phylocsfpp_b200/lib/libphylocsf_b200_$NAME.so
