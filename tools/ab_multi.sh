#!/bin/bash
# The stand-in translation units for the transform sizes that are not being compared live in tools/ab/ (kern_stub*.cu).
# tools/ab_multi.sh name1 "flags1" name2 "flags2" ...   -> builds distantspeechrecognition-mirror_b200/libbtkb200_<name>.so
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  make -C distantspeechrecognition-mirror_b200/csrc -j8 VPATH=../../tools/ab OBJDIR=../build_$n TARGET=../libbtkb200_$n.so EXTRA="-I. $f" SRCS="${AB_SRCS:-kern_m256.cu kern_stub.cu kern_misc.cu kern_cov_tc.cu kern_postfilter.cu kern_design.cu capi.cu}" >/dev/null 2>&1 || echo "build $n failed"
done
