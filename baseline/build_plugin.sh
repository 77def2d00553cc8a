#!/usr/bin/env bash
# Builds plugin/hoomd_b200_plugin/_b200_plugin*.so: a HOOMD-blue component (pybind11 module) compiled against
# the headers of the reference tree that baseline/build_ref.sh configured under $HOOMD_REF_WORK, with the
# flags of its own example plugin (build.ninja: hoomd/example_plugins/pair_plugin), linked against the
# installed _hoomd module in baseline/_ref. Run baseline/build_ref.sh first.   Usage: baseline/build_plugin.sh
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
WORK="${HOOMD_REF_WORK:-/tmp/hoomd_ref_build}"
REF="$HERE/_ref/hoomd"
[ -f "$WORK/build/hoomd/include/hoomd/HOOMDVersion.h" ] || [ -d "$WORK/build/hoomd/include" ] || { echo "no configured reference tree at $WORK (run baseline/build_ref.sh)" >&2; exit 2; }
[ -d "$REF" ] || { echo "no reference install at $REF" >&2; exit 2; }
EXT="$(python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
PBINC="$(python -c 'import pybind11; print(pybind11.get_include())')"
OUT="$ROOT/plugin/hoomd_b200_plugin/_b200_plugin$EXT"
g++ -O3 -DNDEBUG -std=gnu++17 -fPIC -fvisibility=hidden -shared \
    -DEIGEN_MPL2_ONLY -DHOOMD_LONGREAL_SIZE=64 -DHOOMD_SHORTREAL_SIZE=32 -D_REENTRANT \
    -Wall -Wno-sign-conversion -Wno-unknown-pragmas -Wno-deprecated-declarations -Wno-unused-result \
    -I"$WORK/src" -I"$WORK/build/hoomd/include" -isystem "$PBINC" -isystem "$PYINC" \
    "$ROOT/plugin/hoomd_b200_plugin/module.cc" -o "$OUT" \
    -L"$REF" -l:_hoomd$EXT -ldl \
    -Wl,-rpath,'$ORIGIN/../../baseline/_ref/hoomd'
echo "built $OUT"
