#!/bin/bash
# tools/ab_build.sh <git-rev> <name>: builds libkhg_b200.so of an earlier revision into
# tools/ab/<name>.so so that two builds can be timed on the SAME GPU box in one gpurun call
# (boxes differ by >10 % under the power cap):  KHG_B200_LIB=tools/ab/<name>.so python tools/k1_modes.py ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
rev=$1; name=$2
tmp=$(mktemp -d)
git -C "$ROOT" archive "$rev" kaldi-hmm-gmm_b200/csrc include | tar -x -C "$tmp"
cd "$tmp/kaldi-hmm-gmm_b200/csrc"
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    -I"$tmp/include" -I. --expt-relaxed-constexpr -c $f -o ${f%.cu}.o 2>/dev/null &
done
wait
mkdir -p "$ROOT/tools/ab"
/usr/local/cuda/bin/nvcc -shared -o "$ROOT/tools/ab/$name.so" *.o -cudart shared
rm -rf "$tmp"
echo "built tools/ab/$name.so from $rev"
