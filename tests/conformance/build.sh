#!/bin/sh
# Builds the conformance driver from the reference's UNMODIFIED wrapper sources (read in place under
# $REF, default /root/reference; nothing is copied) against this repository's acados-compatible
# headers and libraries.  Output only into tests/conformance/_build/ (git-ignored, travels to the GPU box).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${REF:-/root/reference}
OUT="$HERE/_build"
mkdir -p "$OUT"
g++ -std=c++14 -O2 -Wall -Wextra \
    -I "$ROOT/include" -I "$HERE/stubs" -I "$REF/include" \
    "$HERE/driver.cpp" \
    "$REF/src/nmpc_nav_control/NMPCNavControl.cpp" \
    "$REF/src/nmpc_nav_control/NMPCNavControlDiff.cpp" \
    "$REF/src/nmpc_nav_control/NMPCNavControlOmni4.cpp" \
    "$REF/src/nmpc_nav_control/NMPCNavControlTric.cpp" \
    -L "$ROOT/nmpc_nav_control_b200" \
    -lacados_ocp_solver_diff2amr -lacados_ocp_solver_omni4amr -lacados_ocp_solver_tric3amr -lacados -lnmpc_b200 \
    -Wl,-rpath,'$ORIGIN/../../../nmpc_nav_control_b200' \
    -o "$OUT/conformance_driver"
echo "built $OUT/conformance_driver"
