#!/bin/bash
# Builds libnerf_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libnerf_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr ${EXTRA_NVCC_FLAGS:-})
OBJS=()
PIDS=()
for f in ops mlp_fp32 ctx mlp_tc mlp_tc_experimental mlp_tc_bwd bn_train gemm_tc; do
  src="$HERE/$f.cu"; obj="$HERE/$f.o"
  stale=0
  [[ ! -f "$obj" || "$src" -nt "$obj" || "$HERE/../../include/nerf_b200.h" -nt "$obj" || "$HERE/../../include/nerf_b200_debug.h" -nt "$obj" ]] && stale=1
  for hdr in "$HERE"/*.cuh; do [[ "$hdr" -nt "$obj" ]] && stale=1; done
  if [[ $stale -eq 1 ]]; then
    ( "$NVCC" "${FLAGS[@]}" ${PTXAS_V:+-Xptxas -v} -c "$src" -o "$obj" || { rm -f "$obj"; exit 1; } ) &
    PIDS+=($!)
  fi
  OBJS+=("$obj")
done
for pid in "${PIDS[@]:-}"; do [[ -n "$pid" ]] && wait "$pid"; done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}" -lcudart -ldl
echo "built $OUT"
