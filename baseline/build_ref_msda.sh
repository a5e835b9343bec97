#!/bin/bash
# Compiles the reference's MSDeformAttn CUDA kernels for sm_100a into baseline/_ref/libref_msda.so (git-ignored; it
# travels to the GPU box with the snapshot).  Needs the reference checkout (this container only); sources stay where they are.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PDB_REFERENCE:-/root/reference}/part_distillation/modeling/pixel_decoder/ops/src"
[ -f "$REF/cuda/ms_deform_im2col_cuda.cuh" ] || { echo "reference checkout not found at $REF" >&2; exit 3; }
mkdir -p "$HERE/_ref"
${NVCC:-/usr/local/cuda/bin/nvcc} -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
    -I "$HERE/ref_msda/shim" -I "$REF" "$HERE/ref_msda/ref_msda.cu" -o "$HERE/_ref/libref_msda.so"
echo "$HERE/_ref/libref_msda.so"
