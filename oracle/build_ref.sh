#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle). Compiles the reference's OWN native kernels from where they lie under
# /root/reference into oracle/_ref/libjvector.so. Flags restate jvector-native/src/main/native/meson.build:28-110
# (three ISA builds of jvector_simd_kernels.cpp + two tier TUs + dispatcher + hwy/abort.cc). No reference source
# is copied into this repo; outputs go only to oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${JVECTOR_REFERENCE:-/root/reference}"
N="$REF/jvector-native/src/main/native"
H="$N/third_party/highway"
OUT="$HERE/_ref"
if [ ! -d "$N/src" ]; then
  if [ -f "$OUT/libjvector.so" ]; then echo "reference absent; keeping prebuilt $OUT/libjvector.so"; exit 0; fi
  echo "reference sources not found at $N and no prebuilt oracle/_ref/libjvector.so" >&2; exit 3
fi
if [ -f "$OUT/libjvector.so" ] && [ "$OUT/libjvector.so" -nt "$N/src/jvector_simd_kernels.cpp" ] && [ "${FORCE:-0}" != 1 ]; then
  echo "up to date: $OUT/libjvector.so"; exit 0
fi
mkdir -p "$OUT/obj"
F="-std=c++17 -O2 -fPIC -fvisibility=hidden -I$H"
O="$OUT/obj"
g++ $F -march=skylake-avx512 -DHWY_COMPILE_ONLY_STATIC -DJV_REQUIRE_HWY_AVX3 -DJV_ISA=AVX3 -c "$N/src/jvector_simd_kernels.cpp" -o "$O/k_avx3.o" &
g++ $F -march=haswell -maes -DHWY_COMPILE_ONLY_STATIC -DJV_REQUIRE_HWY_AVX2 -DJV_ISA=AVX2 -c "$N/src/jvector_simd_kernels.cpp" -o "$O/k_avx2.o" &
g++ $F -msse4.2 -mpclmul -maes -DHWY_COMPILE_ONLY_STATIC -DJV_REQUIRE_HWY_SCALAR -DJV_ISA=SSE42 -c "$N/src/jvector_simd_kernels.cpp" -o "$O/k_sse.o" &
g++ $F -march=icelake-server -DHWY_COMPILE_ONLY_STATIC -DJV_REQUIRE_HWY_AVX3_DL -c "$N/src/jvector_avx3_dl_kernels.cpp" -o "$O/k_dl.o" &
g++ $F -march=sapphirerapids -DHWY_COMPILE_ONLY_STATIC -DJV_REQUIRE_HWY_AVX3_SPR -c "$N/src/jvector_avx3_spr_kernels.cpp" -o "$O/k_spr.o" &
g++ $F -I"$N/src" -DJVECTOR_BUILD -c "$N/src/jvector_simd.cpp" -o "$O/simd.o" &
g++ $F -c "$H/hwy/abort.cc" -o "$O/abort.o" &
wait
g++ -shared -o "$OUT/libjvector.so" "$O/simd.o" "$O/abort.o" "$O/k_avx3.o" "$O/k_avx2.o" "$O/k_sse.o" "$O/k_dl.o" "$O/k_spr.o"
rm -rf "$O"
echo "built $OUT/libjvector.so"
