#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds the UNMODIFIED reference tracer into oracle/_ref/ref_render.
#
# Compiles /root/reference/{3DElement,Basic3DObject,Model,Scene,RayTracer}.cpp where they lie
#  with the GL/Win32 shim in oracle/shim and links them with
# tools/render_main.cpp (-DRT_ARM_REFERENCE, oracle/render_taps.h).  Model.cpp needs five
# MSVC-only spellings rewritten; that is done on a throw-away copy in a mktemp dir which is
# deleted again -- no reference source is ever written into this repository.  Outputs go to
# oracle/_ref/ only (git-ignored; travels to the GPU box with gpurun).
# Flags: -O2 -mavx2 -mfma -ffp-contract=off  ==  "the contraction-off g++ build" that
# SURVEY.md 8c defines as the project's oracle.
set -euo pipefail
REF=${RT_REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -f "$REF/RayTracer.cpp" ]; then
	echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)" >&2
	exit 0
fi
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
sed -E \
	-e 's/\.swap\(vector<[^;]*>\(\)\);/.clear();/' \
	-e 's/\.m128_f32\[/[/g' -e 's/\.m256_f32\[/[/g' \
	-e 's/([A-Za-z_]+)\.m256i_i32\[([a-z])\]/((const int*)\&\1)[\2]/g' \
	-e 's/\*\(__m256i\*\)&_mm256_cmp_ps\(ansmin, ansmax, _CMP_LE_OS\)/_mm256_castps_si256(_mm256_cmp_ps(ansmin, ansmax, _CMP_LE_OS))/' \
	"$REF/Model.cpp" > "$TMP/Model.cpp"
CXX=${CXX:-g++}
FLAGS="-std=c++14 -fpermissive -w -O2 -mavx2 -mfma -ffp-contract=off -I $HERE/shim -I $REF -include $HERE/shim/pre.h"
$CXX $FLAGS -fkeep-inline-functions -c "$REF/3DElement.cpp" -o "$TMP/3DElement.o" &
$CXX $FLAGS -c "$REF/Basic3DObject.cpp" -o "$TMP/Basic3DObject.o" &
$CXX $FLAGS -c "$TMP/Model.cpp" -o "$TMP/Model.o" &
$CXX $FLAGS -c "$REF/Scene.cpp" -o "$TMP/Scene.o" &
$CXX $FLAGS -c "$REF/RayTracer.cpp" -o "$TMP/RayTracer.o" &
$CXX $FLAGS -fno-access-control -DRT_ARM_REFERENCE -I "$HERE" -c "$HERE/../tools/render_main.cpp" -o "$TMP/render_main.o" &
wait
$CXX -o "$OUT/ref_render.new" "$TMP"/*.o -lpthread
mv -f "$OUT/ref_render.new" "$OUT/ref_render"   # atomic: a running ref_render keeps its old image
echo "built $OUT/ref_render"
