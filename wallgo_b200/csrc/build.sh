#!/usr/bin/env bash
# Builds wallgo_b200/libwallgo_b200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libwallgo_b200.so"
OBJ="$HERE/../_obj"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH ${WG_NVCC_EXTRA:-}"
pids=()
"$NVCC" $COMMON -c "$HERE/wg_gemm.cu" -o "$OBJ/wg_gemm.o" & pids+=($!)
"$NVCC" $COMMON -c "$HERE/wg_lu.cu" -o "$OBJ/wg_lu.o" & pids+=($!)
"$NVCC" $COMMON -fmad=false -c "$HERE/wg_boltzmann.cu" -o "$OBJ/wg_boltzmann.o" & pids+=($!)
"$NVCC" $COMMON -c "$HERE/wg_capi.cu" -o "$OBJ/wg_capi.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" $ARCH -shared -o "$OUT" "$OBJ/wg_gemm.o" "$OBJ/wg_lu.o" "$OBJ/wg_boltzmann.o" "$OBJ/wg_capi.o" -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
