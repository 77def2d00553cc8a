#!/usr/bin/env bash
# Build the UNMODIFIED-on-the-hot-path reference (HOOMD-blue v4.8.2, CPU, fp64, no MPI) into the
# git-ignored baseline/_ref/ so that
#   * tests/golden/make_golden.py can regenerate every fixture,
#   * bench.py --impl reference / cpu_baseline can time the real hoomd.device.CPU (kind "reference"),
#   * -m gpu tests can run the reference at the configured sizes when the build travelled to the box.
# Recipe = SURVEY.md section 8(c).  The only edits to the copy under $WORK remove the two users of
# Eigen (absent from this image), neither of which is on the pair-force path:
#   hoomd/extern/BVLSSolver.cc (+ its include in LoadBalancer.cc, MPI-only code) and
#   hoomd/md/ForceDistanceConstraint.{cc,h} (+ its export in module-md.cc).
# /root/reference itself is never written.  Usage:  baseline/build_ref.sh [jobs]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${HOOMD_REFERENCE:-/root/reference}"
WORK="${HOOMD_REF_WORK:-/tmp/hoomd_ref_build}"
OUT="$HERE/_ref"
JOBS="${1:-$(nproc)}"

[ -d "$REF/hoomd" ] || { echo "reference tree not found at $REF" >&2; exit 2; }
rm -rf "$WORK"; mkdir -p "$WORK"
cp -r "$REF" "$WORK/src"
chmod -R u+w "$WORK/src"
S="$WORK/src"

# (1)(2) BVLSSolver / LoadBalancer
sed -i '/extern\/BVLSSolver.cc/d' "$S/hoomd/CMakeLists.txt"
sed -i 's/ Eigen3::Eigen)/)/' "$S/hoomd/CMakeLists.txt"
python - "$S/hoomd/LoadBalancer.cc" <<'EOF'
import sys
p = sys.argv[1]
t = open(p).read()
t = t.replace('#include "hoomd/extern/BVLSSolver.h"\n#include <Eigen/Dense>\n',
              '#ifdef ENABLE_MPI\n#include "hoomd/extern/BVLSSolver.h"\n#include <Eigen/Dense>\n#endif\n')
open(p, 'w').write(t)
EOF
# (3)(4) ForceDistanceConstraint
sed -i '/^ *ForceDistanceConstraint\.cc$/d; /^ *ForceDistanceConstraint\.h$/d' "$S/hoomd/md/CMakeLists.txt"
sed -i '/^void export_ForceDistanceConstraint(pybind11::module& m);$/d; /^ *export_ForceDistanceConstraint(m);$/d' "$S/hoomd/md/module-md.cc"

# Eigen stub: find_package(Eigen3 3.2 CONFIG REQUIRED) is unconditional
mkdir -p "$WORK/eigen_stub/cmake" "$WORK/eigen_stub/include/Eigen"
: > "$WORK/eigen_stub/include/Eigen/Dense"
cat > "$WORK/eigen_stub/cmake/Eigen3Config.cmake" <<EOF
set(Eigen3_FOUND TRUE)
set(Eigen3_VERSION 3.4.0)
set(EIGEN3_INCLUDE_DIR "$WORK/eigen_stub/include")
if(NOT TARGET Eigen3::Eigen)
  add_library(Eigen3::Eigen INTERFACE IMPORTED)
  set_target_properties(Eigen3::Eigen PROPERTIES INTERFACE_INCLUDE_DIRECTORIES "$WORK/eigen_stub/include")
endif()
EOF
cat > "$WORK/eigen_stub/cmake/Eigen3ConfigVersion.cmake" <<'EOF'
set(PACKAGE_VERSION 3.4.0)
set(PACKAGE_VERSION_COMPATIBLE TRUE)
EOF

rm -rf "$OUT"; mkdir -p "$OUT"
cmake -S "$S" -B "$WORK/build" -G Ninja \
  -DCMAKE_BUILD_TYPE=Release -DENABLE_GPU=OFF -DENABLE_MPI=OFF -DENABLE_TBB=OFF -DENABLE_LLVM=OFF \
  -DBUILD_HPMC=OFF -DBUILD_MPCD=OFF -DBUILD_METAL=OFF -DBUILD_TESTING=OFF \
  -DPLUGINS=example_plugins/pair_plugin \
  -DEigen3_DIR="$WORK/eigen_stub/cmake" \
  -Dpybind11_DIR="$(python -m pybind11 --cmakedir)" \
  -DCMAKE_INSTALL_PREFIX="$OUT" -DPYTHON_SITE_INSTALL_DIR:STRING=hoomd \
  > "$WORK/configure.log" 2>&1 || { tail -40 "$WORK/configure.log"; exit 3; }
ninja -C "$WORK/build" -j "$JOBS" > "$WORK/build.log" 2>&1 || { tail -60 "$WORK/build.log"; exit 4; }
ninja -C "$WORK/build" install > "$WORK/install.log" 2>&1 || { tail -40 "$WORK/install.log"; exit 5; }
# keep the box snapshot small: sources of the tests and headers are not needed at run time
find "$OUT" -name '*.h' -o -name '*.cuh' -o -name '*.hpp' | xargs -r rm -f
PYTHONPATH="$OUT" python - <<'EOF'
import hoomd
print("reference build ok:", hoomd.version.version, hoomd.version.floating_point_precision,
      "md", hoomd.version.md_built, "mpi", hoomd.version.mpi_enabled, "gpu", hoomd.version.gpu_enabled)
EOF
rm -rf "$OUT/lib"
du -sh "$OUT"
