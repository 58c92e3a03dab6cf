#!/usr/bin/env bash
# Build oracle/_ref/libnawsod_ref.so: the reference's own CPU operators for the hot path,
# compiled UNMODIFIED from the sources where they lie under /root/reference (never copied),
# against oracle/c2shim (a stand-in for the Caffe2 v1.3.0 operator API, which is not
# vendored and not installable here).  Flags follow the reference's CMakeLists.txt:24
# (-std=c++11 -O2).  Outputs only under oracle/_ref/ (git-ignored, travels via gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NAWSOD_REFERENCE:-/root/reference}"
OPS="$REF/detectron/ops"
if [ ! -d "$OPS" ]; then
  echo "build_ref: $OPS not present (GPU box): keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$HERE/_ref"
g++ -std=c++11 -O2 -fPIC -shared -w \
  -I"$HERE/c2shim" -I"$OPS" \
  "$OPS/roi_feature_boost_op.cc" \
  "$OPS/cross_entropy_wsl_op.cc" \
  "$OPS/acm_weightdecay_momentum_sgd_op.cc" \
  "$HERE/ref_driver.cc" \
  -o "$HERE/_ref/libnawsod_ref.so"
echo "built $HERE/_ref/libnawsod_ref.so"
