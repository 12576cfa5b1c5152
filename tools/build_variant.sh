#!/bin/bash
# A/B builds: compile the named .cu files with extra -D flags and link them with the objects of the
# current build into tools/scratch/libs/libct_<name>.so (load it with CT_B200_LIB=...).
# usage: tools/build_variant.sh <name> "<-D flags>" [file.cu ...]   (default file: ct_idt.cu)
set -e
cd "$(dirname "$0")/.."
name=$1; defs=$2; shift 2 || true
files=${@:-ct_idt.cu}
C=color-transfer_b200/csrc
mkdir -p tools/scratch/libs/obj_$name
objs=""
for f in ct_api ct_linear ct_idt ct_u8 ct_regrain ct_metrics ct_distort; do
  if [[ " $files " == *" $f.cu "* ]]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr $defs -c $C/$f.cu -o tools/scratch/libs/obj_$name/$f.o
    objs="$objs tools/scratch/libs/obj_$name/$f.o"
  else
    objs="$objs $C/$f.o"
  fi
done
nvcc -shared -o tools/scratch/libs/libct_$name.so $objs -gencode arch=compute_100a,code=sm_100a
rm -rf tools/scratch/libs/obj_$name
echo tools/scratch/libs/libct_$name.so
