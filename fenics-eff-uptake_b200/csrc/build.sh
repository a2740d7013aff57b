#!/usr/bin/env bash
# Build libsulcusfem.so for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../sulcusfem/libsulcusfem.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=(sfem_vector.cu sfem_spmv.cu sfem_spmv_staged.cu sfem_assembly.cu sfem_mg.cu sfem_krylov.cu sfem_functionals.cu sfem_api.cu)
cd "$here"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
  ${SFEM_PTXAS_V:+-Xptxas -v} -o "$out" "${srcs[@]}"
echo "built $out"
