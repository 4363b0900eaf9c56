#!/bin/sh
# build_ref.sh <reference root> — compiles the reference's OWN shader sources as C++.
# The .glsl files are read where they lie (never copied): a sed stream strips `#version`, gives
# bare float literals an `f` suffix (GLSL literals are float32, C++ ones would be double) and
# turns `inout float x` into a C++ reference; the result is piped straight into g++ between a
# prologue that opens a namespace and the keyword macros.  Outputs only into oracle/_ref/.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_ref"
CXX=/usr/bin/g++
FLAGS="-O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -w -I$HERE"
mkdir -p "$OUT"
[ -d "$REF/res" ] || { echo "no reference tree at $REF" >&2; exit 1; }
OBJS=""
for sh in conetrace_frag first_voxelize second_voxelize sun_frag billboard_vert_instanced; do
  {
    echo '#include "glsl_shim.hpp"'
    echo "namespace $sh { using namespace glsl;"
    echo 'vec4 gl_FragCoord; float gl_FragDepth; vec4 gl_Position; bool gl_Discarded;'
    echo '#include "glsl_keywords.hpp"'
    sed -E -e '/^[[:space:]]*#version/d' \
           -e 's/\binout[[:space:]]+float[[:space:]]+/float \&/g' \
           -e 's/([0-9]+\.[0-9]*)([^0-9f.]|$)/\1f\2/g' "$REF/res/$sh.glsl"
    echo
    echo '}'
  } | $CXX $FLAGS -x c++ -c - -o "$OUT/$sh.o"
  OBJS="$OBJS $OUT/$sh.o"
done
$CXX $FLAGS -c "$HERE/ref_api.cpp" -o "$OUT/ref_api.o"
$CXX -shared -o "$OUT/libref_glsl.so" $OBJS "$OUT/ref_api.o" -L"$HERE/.." -loracle -Wl,-rpath,'$ORIGIN/..'
rm -f $OBJS "$OUT/ref_api.o"
echo "built $OUT/libref_glsl.so"
