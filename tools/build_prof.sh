#!/bin/bash
# profile build (phase clocks, bgmm_sweep_stats.phase_cycles) of the D=2 and D=16 engines into
# pybgmm_b200/lib/libbgmm_b200_prof.so; use with BGMM_B200_LIB=pybgmm_b200/lib/libbgmm_b200_prof.so
set -e
cd "$(dirname "$0")/.."
B=pybgmm_b200/build/prof; mkdir -p $B
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -diag-suppress 128 -DBGMM_PROFILE"
nvcc $F -DBGMM_HAVE_D16 -DBGMM_HAVE_D2 -c pybgmm_b200/csrc/bgmm_engine.cu -o $B/engine.o &
nvcc $F -c pybgmm_b200/build/inst/inst_full_16.cu -o $B/full16.o &
nvcc $F -c pybgmm_b200/build/inst/inst_diag_16.cu -o $B/diag16.o &
nvcc $F -c pybgmm_b200/build/inst/inst_full_2.cu -o $B/full2.o &
nvcc $F -c pybgmm_b200/build/inst/inst_diag_2.cu -o $B/diag2.o &
nvcc $F -c pybgmm_b200/build/inst/inst_fixed_16.cu -o $B/fixed16.o &
nvcc $F -c pybgmm_b200/build/inst/inst_fixed_2.cu -o $B/fixed2.o &
g++ -O2 -fPIC -c pybgmm_b200/csrc/mt19937.cc -o $B/mt.o &
wait
for o in engine full16 diag16 full2 diag2 fixed16 fixed2 mt; do test $B/$o.o -nt pybgmm_b200/csrc/bgmm_seq.cuh -a $B/$o.o -nt pybgmm_b200/csrc/bgmm_clu.cuh || { echo "stale or missing $o.o: compile failed"; exit 1; }; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $B/lib.tmp $B/engine.o $B/full16.o $B/diag16.o $B/full2.o $B/diag2.o $B/fixed16.o $B/fixed2.o $B/mt.o && mv $B/lib.tmp pybgmm_b200/lib/libbgmm_b200_prof.so
echo built pybgmm_b200/lib/libbgmm_b200_prof.so
