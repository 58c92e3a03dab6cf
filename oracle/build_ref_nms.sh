#!/usr/bin/env bash
# Build oracle/_ref/cython_nms*.so: the reference's own greedy NMS (detectron/utils/cython_nms.pyx),
# cythonized UNMODIFIED from the source where it lies under /root/reference (never copied).  The
# reference was written against numpy 1.x: `np.int_t` was removed from numpy 2's .pxd, so the build
# overlays a copy of THIS numpy's own __init__.pxd plus that one typedef (a build-time shim written to
# oracle/_ref/pxd/, like oracle/c2shim for the Caffe2 API); callers set `numpy.int = int` before use.
# Outputs only under oracle/_ref/ (git-ignored, travels to the GPU box via gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NAWSOD_REFERENCE:-/root/reference}"
PYX="$REF/detectron/utils/cython_nms.pyx"
if [ ! -f "$PYX" ]; then
  echo "build_ref_nms: $PYX not present (GPU box): keeping prebuilt oracle/_ref" >&2
  exit 0
fi
OUT="$HERE/_ref"
NP="$(python -c 'import numpy,os;print(os.path.dirname(numpy.__file__))')"
EXT="$(python -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
PYINC="$(python -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
mkdir -p "$OUT/pxd/numpy"
cp "$NP/__init__.cython-30.pxd" "$OUT/pxd/numpy/__init__.pxd"
printf '\n# numpy-1.x name the reference still uses (removed in numpy 2)\nctypedef npy_long       int_t\n' >> "$OUT/pxd/numpy/__init__.pxd"
python -m cython -3 --fast-fail -I "$OUT/pxd" -o "$OUT/cython_nms.c" "$PYX"
gcc -O2 -fPIC -shared -w -I"$PYINC" -I"$NP/_core/include" "$OUT/cython_nms.c" -o "$OUT/cython_nms$EXT"
rm -f "$OUT/cython_nms.c"
echo "built $OUT/cython_nms$EXT"
