#!/usr/bin/env bash
# Build oracle/_build/liboracle_roi_pool.so (the C restatement; gcc + OpenMP).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$HERE/_build"
gcc -O2 -std=c11 -fPIC -shared -fopenmp -ffp-contract=off "$HERE/roi_pool_ref.c" -lm -o "$HERE/_build/liboracle_roi_pool.so"
echo "built $HERE/_build/liboracle_roi_pool.so"
