#!/usr/bin/env bash
# Build libsulcusfem.so for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
# Translation units are compiled in parallel into build/ and linked into sulcusfem/libsulcusfem.so.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../sulcusfem/libsulcusfem.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=(sfem_vector.cu sfem_spmv.cu sfem_spmv_staged.cu sfem_spmv_sell.cu sfem_assembly.cu sfem_mg.cu sfem_mg_tail.cu sfem_krylov.cu sfem_batch.cu sfem_stokes.cu
      sfem_functionals.cu sfem_points.cu sfem_api.cu sfem_dist.cu)
flags=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC ${SFEM_PTXAS_V:+-Xptxas -v})
cd "$here"
mkdir -p build
pids=()
objs=()
for s in "${srcs[@]}"; do
  [ -f "$s" ] || continue
  o="build/${s%.cu}.o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o" -print -quit)" ] \
     || [ ../../include/sulcusfem.h -nt "$o" ]; then
    "$NVCC" "${flags[@]}" -c -o "$o" "$s" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out" "${objs[@]}"
echo "built $out"
