#!/bin/bash
# developer build of the D=2 and D=16 engines with extra flags into pybgmm_b200/lib/libbgmm_b200_<name>.so
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"; use with BGMM_B200_LIB=pybgmm_b200/lib/libbgmm_b200_<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2
B=pybgmm_b200/build/var_$NAME; mkdir -p $B; rm -f $B/*.o
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -diag-suppress 128 $EXTRA"
nvcc $F -DBGMM_HAVE_D16 -DBGMM_HAVE_D2 -c pybgmm_b200/csrc/bgmm_engine.cu -o $B/engine.o &
for u in full_16 diag_16 fixed_16 full_2 diag_2 fixed_2; do nvcc $F -c pybgmm_b200/build/inst/inst_$u.cu -o $B/$u.o & done
g++ -O2 -fPIC -c pybgmm_b200/csrc/mt19937.cc -o $B/mt.o &
wait
for o in engine full_16 diag_16 fixed_16 full_2 diag_2 fixed_2 mt; do test -f $B/$o.o || { echo "missing $o.o: compile failed"; exit 1; }; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $B/lib.tmp $B/*.o && mv $B/lib.tmp pybgmm_b200/lib/libbgmm_b200_$NAME.so
echo built pybgmm_b200/lib/libbgmm_b200_$NAME.so
