#!/bin/bash
# tools/ab_multi.sh name1 "flags1" name2 "flags2" ...   -> builds distantspeechrecognition-mirror_b200/libbtkb200_<name>.so
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  make -C distantspeechrecognition-mirror_b200/csrc -j8 OBJDIR=../build_$n TARGET=../libbtkb200_$n.so EXTRA="$f" SRCS="${AB_SRCS:-kern_m256.cu kern_stub.cu kern_misc.cu kern_cov_tc.cu kern_postfilter.cu kern_design.cu capi.cu}" >/dev/null 2>&1 || echo "build $n failed"
done
