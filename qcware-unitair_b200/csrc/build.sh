#!/usr/bin/env bash
# Build libunitair_b200.so (pure C ABI, sm_100a only) in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../unitair_b200/lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr
       -Xcompiler -fPIC ${UA_NVCC_EXTRA:-})
objs=()
pids=()
for f in ua_api ua_gate ua_phase ua_reduce ua_grad ua_tile ua_cluster ua_tc5 ua_permute ua_diag ua_sample; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/$f.o" &
  pids+=($!)
  objs+=("$HERE/$f.o")
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libunitair_b200.so" "${objs[@]}"
echo "built $OUT/libunitair_b200.so"
