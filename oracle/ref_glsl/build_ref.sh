#!/bin/sh
# build_ref.sh <reference root> — compiles the reference's OWN shader sources, and its host code on the hot path, as C++.
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
# The paper variant: the same first_voxelize.glsl with its commented-out interior march (res/first_voxelize.glsl:54-58)
# switched back on by stripping the "// " in front of those lines, as a second namespace.
{
  echo '#include "glsl_shim.hpp"'
  echo "namespace first_voxelize_paper { using namespace glsl;"
  echo 'vec4 gl_FragCoord; float gl_FragDepth; vec4 gl_Position; bool gl_Discarded;'
  echo '#include "glsl_keywords.hpp"'
  sed -E -e '/^[[:space:]]*#version/d' \
         -e '/Write to volume in spherical shape/,/Write nearest voxel position/s|^([[:space:]]*)// |\1|' \
         -e 's/\binout[[:space:]]+float[[:space:]]+/float \&/g' \
         -e 's/([0-9]+\.[0-9]*)([^0-9f.]|$)/\1f\2/g' "$REF/res/first_voxelize.glsl"
  echo
  echo '}'
} | $CXX $FLAGS -x c++ -c - -o "$OUT/first_voxelize_paper.o"
OBJS="$OBJS $OUT/first_voxelize_paper.o"

# The reference's HOST code on the hot path, compiled from where it lies (src/), against stub third-party headers
# (host_shim/: GLM restated from its published definitions, GLFW / glad names only).
HOSTFLAGS="$FLAGS -I$HERE/host_shim -I$REF/src"
$CXX $HOSTFLAGS -Dprotected=public -c "$REF/src/Camera.cpp" -o "$OUT/host_camera.o"
{   # Sun's static members with the reference's initial values (src/main.cpp:37-46)
  echo '#include "Sun.hpp"'
  grep -E '^(glm::(vec3|mat4)|float) Sun::' "$REF/src/main.cpp"
} | $CXX $HOSTFLAGS -x c++ -c - -o "$OUT/host_sun.o"
{   # CloudVolume: constructor head, addCloudBoard + sortBoards, update, get3DIndices .. resetBillboards
  echo '#include "CloudVolume.hpp"'
  echo '#include "Util.hpp"'
  sed -n '/^CloudVolume::CloudVolume/,/this->levels = mips;/p' "$REF/src/CloudVolume.cpp"; echo '}'
  sed -n '/^\/\* Add a billboard \*\//,/^void CloudVolume::update/p' "$REF/src/CloudVolume.cpp" | sed '$d'
  sed -n '/^void CloudVolume::update/,/^}/p' "$REF/src/CloudVolume.cpp"
  sed -n '/^\/\/ Assume 4 bytes per voxel/,/^void CloudVolume::uploadBillboards/p' "$REF/src/CloudVolume.cpp" | sed '$d'
} | $CXX $HOSTFLAGS -x c++ -c - -o "$OUT/host_volume.o"
{   # initNoiseMap: its helpers verbatim, and the normal loop as the body of a free function
  echo '#include "glm/glm.hpp"'
  sed -n '/^int getIndex/,/^void ConeTraceShader::initNoiseMap/p' "$REF/src/Shaders/ConeTraceShader.cpp" | sed '$d'
  echo 'void ref_noise_normals_impl(CHAR4 *pData, int dimension) {'
  sed -n '/Generate normals from the density gradient/,/glGenTextures(1, &noiseMapId)/p' "$REF/src/Shaders/ConeTraceShader.cpp" | sed '$d'
  echo '}'
} | $CXX $HOSTFLAGS -x c++ -c - -o "$OUT/host_noise.o"
$CXX $HOSTFLAGS -c "$HERE/ref_host.cpp" -o "$OUT/ref_host.o"
OBJS="$OBJS $OUT/host_camera.o $OUT/host_sun.o $OUT/host_volume.o $OUT/host_noise.o $OUT/ref_host.o"

$CXX $FLAGS -c "$HERE/ref_api.cpp" -o "$OUT/ref_api.o"
$CXX -shared -o "$OUT/libref_glsl.so" $OBJS "$OUT/ref_api.o" -L"$HERE/.." -loracle -Wl,-rpath,'$ORIGIN/..'
rm -f $OBJS "$OUT/ref_api.o"
echo "built $OUT/libref_glsl.so"
