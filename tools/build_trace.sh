#!/bin/bash
# builds the PCSF_TC5_TRACE variant of the library (per-step clock64 timestamps of one CTA, tools/tc5_trace_run.py)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static -DPCSF_TC5_TRACE \
  -o phylocsfpp_b200/lib/libphylocsf_b200_trace.so phylocsfpp_b200/csrc/pcsf_capi.cu 2>&1 | grep -v "warning\|^$\|\^\|detected during\|Remark\|q == 0" | head
