#!/bin/sh
# TEST INFRASTRUCTURE ONLY. Compiles the reference's own WRMF headers, where they lie under
# /root/reference (read-only, never copied), against oracle/mini_arma into oracle/_ref/.
# Flags follow the reference build (src/Makevars.in:1): -DARMA_32BIT_WORD -DARMA_NO_DEBUG,
# OpenMP on; RSPARSE_R_PKG is left undefined so wrmf.hpp:1-5 picks <armadillo> (ours).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${RSPARSE_REFERENCE:-/root/reference}"
[ -d "$REF/inst/include" ] || { echo "reference tree not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$HERE/_ref"
CXXFLAGS="-std=c++17 -O2 -march=x86-64-v2 -fopenmp -fPIC -w -DARMA_32BIT_WORD -DARMA_NO_DEBUG"
INC="-I $HERE/mini_arma -I $REF/inst/include"
g++ $CXXFLAGS $INC -DREF_IMPLICIT -c "$HERE/ref_shim.cpp" -o "$HERE/_ref/ref_implicit.o"
g++ $CXXFLAGS $INC -DREF_EXPLICIT -c "$HERE/ref_shim.cpp" -o "$HERE/_ref/ref_explicit.o"
g++ -shared -fopenmp "$HERE/_ref/ref_implicit.o" "$HERE/_ref/ref_explicit.o" -o "$HERE/_ref/libref_wrmf.so"
echo "built $HERE/_ref/libref_wrmf.so"
# top_product (src/matrix_top_product.cpp:20-102), compiled in place against oracle/mini_rcpp + oracle/mini_arma
g++ $CXXFLAGS -I "$HERE/mini_rcpp" $INC -I "$REF/src" -c "$REF/src/matrix_top_product.cpp" -o "$HERE/_ref/ref_topk_src.o"
g++ $CXXFLAGS -I "$HERE/mini_rcpp" $INC -I "$REF/src" -c "$HERE/ref_topk_shim.cpp" -o "$HERE/_ref/ref_topk_shim.o"
g++ -shared -fopenmp "$HERE/_ref/ref_topk_src.o" "$HERE/_ref/ref_topk_shim.o" -o "$HERE/_ref/libref_topk.so"
echo "built $HERE/_ref/libref_topk.so"
