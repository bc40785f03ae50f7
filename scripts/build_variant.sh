#!/bin/bash
# Build a tuning variant of libdbsgym.so: recompile ONE translation unit with extra -D flags and link it with the objects of
# the regular build (python -c 'import __graft_entry__ as g; g.build()' first).  Select it with DBSGYM_LIB=<path>.
# usage: scripts/build_variant.sh <name> <unit> [nvcc flags...]      e.g.  scripts/build_variant.sh w7 step_f32_warp -DDBSGYM_WARP_ENVS=7
set -e
name=$1; unit=$2; shift 2
cd "$(dirname "$0")/../dbsgym_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o build/${unit}_${name}.o ${unit}.cu
objs=$(ls build/*.o | grep -v "_[a-z0-9]*\.o$" | grep -v "build/${unit}.o" || true)
# regular objects are named <unit>.o; variant objects <unit>_<name>.o
objs=""
for o in build/*.o; do
  b=$(basename $o .o)
  if [ -f "$b.cu" ] && [ "$b" != "$unit" ]; then objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libdbsgym_${name}.so $objs build/${unit}_${name}.o
echo "built dbsgym_b200/csrc/libdbsgym_${name}.so"
