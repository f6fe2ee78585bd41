#!/bin/bash
# Build tuning variants of libfmdemod.so into variants/<name>/libfmdemod.so (kernel experiments).
# usage: tools/build_variants.sh name:"-DFMD_STAGES=4 -DFMD_ATAN_TERMS=6" ...
set -e
cd "$(dirname "$0")/.."
NVCC=/usr/local/cuda/bin/nvcc
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  d=variants/$name; mkdir -p $d
  gcc -O2 -std=gnu99 -fPIC -ffp-contract=off -Iinclude -Idemodulator_b200/csrc -c -o $d/fmd_design.o demodulator_b200/csrc/fmd_design.c
  defs=$(for f in $flags; do case $f in -D*) echo -n "$f ";; esac; done)  # gcc sees the -D flags only
  gcc -O2 -std=gnu99 -fPIC -ffp-contract=off $defs -Iinclude -Idemodulator_b200/csrc -c -o $d/fmd_quality.o demodulator_b200/csrc/fmd_quality.c
  $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -Iinclude -Idemodulator_b200/csrc \
      -c -o $d/fmdemod.o demodulator_b200/csrc/fmdemod.cu &
  $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -Iinclude -Idemodulator_b200/csrc \
      -c -o $d/fmd_quality_host.o demodulator_b200/csrc/fmd_quality_host.cu &
  for grp in 0 1 2 3 4; do
    $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -DFMD_INST_GROUP=$grp -Iinclude -Idemodulator_b200/csrc \
        -c -o $d/fmd_fast_inst$grp.o demodulator_b200/csrc/fmd_fast_inst.cu &
  done
  wait
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --no-undefined -o $d/libfmdemod.so $d/fmdemod.o $d/fmd_fast_inst0.o $d/fmd_fast_inst1.o \
      $d/fmd_fast_inst2.o $d/fmd_fast_inst3.o $d/fmd_fast_inst4.o $d/fmd_quality_host.o $d/fmd_design.o $d/fmd_quality.o -lm
  rm -f $d/*.o
  echo "built $d ($flags)"
done
