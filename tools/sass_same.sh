#!/bin/bash
# Is the device code of the product library the same as at a git revision (default HEAD)?  Builds that revision's CUDA sources into
# build/sass_cmp/ and compares the SASS of both libraries (names and line tables aside).  For changes that must not touch the
# GPU-validated kernels (host-side work, experiments behind compile-time switches): "0 differing lines" = nothing to re-validate.
# usage: tools/sass_same.sh [rev]
set -e
cd "$(dirname "$0")/.."
rev=${1:-HEAD}
d=build/sass_cmp
rm -rf $d && mkdir -p $d/include
for f in $(git ls-tree --name-only $rev rnaseqc_b200/csrc/ | grep -v /host); do git show $rev:$f > $d/$(basename $f); done
git show $rev:include/rsq_b200.h > $d/include/rsq_b200.h
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$d/include -shared -o $d/rev.so $d/rsq_kernels.cu $d/rsq_covstats.cu $d/rsq_capi.cu -lcudart
strip_names() { /usr/local/cuda/bin/cuobjdump -sass "$1" | grep -v "File\|line\|identifier\|Function :"; }
n=$(diff <(strip_names $d/rev.so) <(strip_names rnaseqc_b200/lib/librsq_b200.so) | wc -l)
echo "$n differing SASS lines between $rev and rnaseqc_b200/lib/librsq_b200.so"
rm -rf $d
[ "$n" = 0 ]
