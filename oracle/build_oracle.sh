#!/usr/bin/env bash
# Build the oracle's C restatements (gcc): oracle/_build/liboracle_roi_pool.so (RoIPoolF, OpenMP)
# and oracle/_build/liboracle_post.so (greedy NMS).  -ffp-contract=off: no fused multiply-adds, the
# reference's arithmetic is operation by operation.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$HERE/_build"
gcc -O2 -std=c11 -fPIC -shared -fopenmp -ffp-contract=off "$HERE/roi_pool_ref.c" -lm -o "$HERE/_build/liboracle_roi_pool.so"
echo "built $HERE/_build/liboracle_roi_pool.so"
gcc -O2 -std=c11 -fPIC -shared -ffp-contract=off "$HERE/post_ref.c" -lm -o "$HERE/_build/liboracle_post.so"
echo "built $HERE/_build/liboracle_post.so"
