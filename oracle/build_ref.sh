#!/usr/bin/env bash
# build_ref.sh -- compile the reference's OWN implementation of the hot path as a host checker.
#
# TEST INFRASTRUCTURE ONLY.  Reads the sources where they lie under $OVRFSR_REFERENCE
# (default /root/reference, read-only), writes ONLY oracle/_ref/libovrfsr_ref.so
# (git-ignored, not gpurun-ignored so the prebuilt .so travels to the GPU box).
# No reference source is copied into the repository: the extracted kernel lines live
# in a temporary directory that is removed on exit.
#
# What is compiled (reference @ 2146b45):
#   consts_ref.cpp : ffx_a.h + ffx_fsr1.h + NIS_Config.h under A_CPU, as shipped
#   fsr_ref.cpp    : ffx_fsr1.h:239-437 (EASU), :684-769 (RCAS), ffx_a.h:1843-1845, behind an HLSL type shim
#   nis_ref_{scaler,sharpen}.cpp : NIS_Scaler.h verbatim (NIS_SCALER=1 / 0) behind an HLSL type shim
#   cas_consts_ref.cpp : src/cas/ffx_a.h + ffx_cas.h under A_CPU, as shipped (CasSetup)
#   cas_ref.cpp    : src/cas/ffx_cas.h:409-893 (CasFilter), src/cas/ffx_a.h:1455-1457, behind the same type shim
#   dds_ref.cpp    : src/postprocess/ScreenGrab11.cpp:72-208 (DDS structs / pixel-format table) and :819-887 (header fill)
# The reference's own build system (Visual Studio + fxc, src/CMakeLists.txt:155-170) is not run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${OVRFSR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
CXX="${CXX:-g++}"
CXXFLAGS="-O2 -ffp-contract=off -fno-fast-math -fPIC -std=c++17 -pthread -w"

if [ ! -f "$REF/src/fsr/ffx_fsr1.h" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi

TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
FSR1="$REF/src/fsr/ffx_fsr1.h"; FFXA="$REF/src/fsr/ffx_a.h"

# anchors: fail loudly if the pinned line numbers drift
sed -n '239p' "$FSR1" | grep -q 'void FsrEasuTapF('  || { echo "anchor 239 moved"; exit 1; }
sed -n '437p' "$FSR1" | grep -q 'pix=min(max4'        || { echo "anchor 437 moved"; exit 1; }
sed -n '684p' "$FSR1" | grep -q 'void FsrRcasF('      || { echo "anchor 684 moved"; exit 1; }
sed -n '769p' "$FSR1" | grep -q 'return;}'            || { echo "anchor 769 moved"; exit 1; }
sed -n '1843p' "$FFXA" | grep -q 'APrxLoRcpF1'        || { echo "anchor 1843 moved"; exit 1; }

HLSL2CPP='s/\b(inout|out) (A[A-Z]+[0-9])\b/\2\&/g'
sed -n '239,437p' "$FSR1" | sed -E "$HLSL2CPP" > "$TMP/easu_lines.inc"
sed -n '684,769p' "$FSR1" | sed -E "$HLSL2CPP" > "$TMP/rcas_lines.inc"
sed -n '1843,1845p' "$FFXA" > "$TMP/ffx_prx.inc"
# NIS_Scaler.h verbatim except the HLSL float literal of NIS_SCALE_FLOAT (unsuffixed = double in C++)
NIS="$REF/src/nis/NIS_Scaler.h"
grep -q '^#define NIS_SCALE_FLOAT 255.0$' "$NIS" || { echo "NIS_SCALE_FLOAT anchor moved"; exit 1; }
sed 's/^#define NIS_SCALE_FLOAT 255.0$/#define NIS_SCALE_FLOAT 255.0f/' "$NIS" > "$TMP/NIS_Scaler_cpp.h"

# legacy CAS path (never dispatched by the mod)
CAS="$REF/src/cas/ffx_cas.h"; CASA="$REF/src/cas/ffx_a.h"
sed -n '409p' "$CAS" | grep -q 'void CasFilter('      || { echo "CAS anchor 409 moved"; exit 1; }
sed -n '893p' "$CAS" | grep -q '^ }$'                  || { echo "CAS anchor 893 moved"; exit 1; }
sed -n '1455p' "$CASA" | grep -q 'APrxLoSqrtF1'        || { echo "CAS anchor 1455 moved"; exit 1; }
sed -n '409,893p' "$CAS" | sed -E "$HLSL2CPP" > "$TMP/cas_lines.inc"
sed -n '1455,1457p' "$CASA" > "$TMP/cas_prx.inc"

# the capture path's DDS container: the reference's own header structs, pixel-format table and fill
GRAB="$REF/src/postprocess/ScreenGrab11.cpp"
sed -n '72p' "$GRAB" | grep -q 'pragma pack(push,1)'          || { echo "DDS anchor 72 moved"; exit 1; }
sed -n '208p' "$GRAB" | grep -q "MAKEFOURCC('D','X','1','0')"  || { echo "DDS anchor 208 moved"; exit 1; }
sed -n '819p' "$GRAB" | grep -q 'MAX_HEADER_SIZE'              || { echo "DDS anchor 819 moved"; exit 1; }
sed -n '887p' "$GRAB" | grep -q '^    }$'                      || { echo "DDS anchor 887 moved"; exit 1; }
sed -n '72,208p' "$GRAB" > "$TMP/dds_structs.inc"
sed -n '819,887p' "$GRAB" > "$TMP/dds_setup.inc"

OBJS=()
for f in consts_ref fsr_ref nis_ref_scaler nis_ref_sharpen cas_consts_ref cas_ref dds_ref; do
  [ -f "$HERE/ref_shim/$f.cpp" ] || continue
  $CXX $CXXFLAGS -I"$REF/src" -I"$TMP" -c "$HERE/ref_shim/$f.cpp" -o "$TMP/$f.o"
  OBJS+=("$TMP/$f.o")
done
$CXX -shared -pthread -o "$OUT/libovrfsr_ref.so" "${OBJS[@]}"
echo "build_ref: wrote $OUT/libovrfsr_ref.so"
