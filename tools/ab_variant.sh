#!/bin/bash
# tools/ab_variant.sh <name> <extra nvcc flags...>: builds the working tree's libkhg_b200.so with
# extra flags (e.g. -DKHG_EPI_GROUPS=4) into tools/ab/<name>.so for same-box A/B timing.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
tmp=$(mktemp -d)
cd "$ROOT/kaldi-hmm-gmm_b200/csrc"
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    -I"$ROOT/include" -I. --expt-relaxed-constexpr "$@" -c $f -o $tmp/${f%.cu}.o 2>/dev/null || echo "COMPILE FAILED: $f" &
done
wait
mkdir -p "$ROOT/tools/ab"
/usr/local/cuda/bin/nvcc -shared -o "$ROOT/tools/ab/$name.so" $tmp/*.o -cudart shared
rm -rf "$tmp"
echo "built tools/ab/$name.so with $*"
