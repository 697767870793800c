#!/usr/bin/env bash
# Builds libintopt_b200.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
#   build.sh            product library  -> ../lib/libintopt_b200.so
#   build.sh profile    profiling build  -> ../lib/libintopt_b200_profile.so (every unit with -DINTOPT_PROFILE plus the
#                                           micro-benchmark / phase-cycle entry points of include/intopt_b200_profile.h)
# Translation units compile in parallel; an object is rebuilt only when a source or header is newer than it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
MODE="${1:-product}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall)
UNITS=(intopt_capi intopt_inst_fwd intopt_inst_fwd_plain intopt_inst_fwd_tmem intopt_inst_bwd intopt_dense)
LIBNAME=libintopt_b200.so
OBJDIR="$HERE/../../build/obj"
if [ "$MODE" = profile ]; then
  FLAGS+=(-DINTOPT_PROFILE)
  UNITS+=(intopt_profile)
  LIBNAME=libintopt_b200_profile.so
  OBJDIR="$HERE/../../build/obj_profile"
fi
# shellcheck disable=SC2206
FLAGS+=(${INTOPT_NVCC_EXTRA:-})
# A/B builds: INTOPT_VARIANT=name (with INTOPT_NVCC_EXTRA="-D...") -> ../lib/libintopt_b200_name.so, objects apart
if [ -n "${INTOPT_VARIANT:-}" ]; then
  LIBNAME="${LIBNAME%.so}_${INTOPT_VARIANT}.so"
  OBJDIR="${OBJDIR}_${INTOPT_VARIANT}"
fi
mkdir -p "$OUT" "$OBJDIR"
newest_dep=$(ls -t "$HERE"/*.cuh "$HERE"/*.hpp "$HERE"/../../include/*.h "$HERE/build.sh" | head -1)
pids=()
names=()
for u in "${UNITS[@]}"; do
  src="$HERE/$u.cu"
  obj="$OBJDIR/$u.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$newest_dep" -nt "$obj" ] || [ -n "${INTOPT_NVCC_EXTRA:-}" ]; then
    "$NVCC" "${FLAGS[@]}" -c -o "$obj" "$src" > "$OBJDIR/$u.log" 2>&1 &
    pids+=($!)
    names+=("$u")
  fi
done
rc=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then
    echo "nvcc failed on ${names[$i]}.cu:" >&2
    cat "$OBJDIR/${names[$i]}.log" >&2
    rc=1
  else
    grep -v '^$' "$OBJDIR/${names[$i]}.log" >&2 || true
  fi
done
[ $rc -eq 0 ] || exit $rc
objs=()
for u in "${UNITS[@]}"; do objs+=("$OBJDIR/$u.o"); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/$LIBNAME" "${objs[@]}"
echo "built $OUT/$LIBNAME"
