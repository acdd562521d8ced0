#!/bin/bash
# A/B two builds of the library on one GPU box: tools/ab.sh "<EXTRA nvcc flags for B>"  (A = default build)
# Builds only kern_m256 + misc + capi for B into build_b/ and alternates bench runs.
set -e
cd "$(dirname "$0")/.."
make -C distantspeechrecognition-mirror_b200/csrc -j8 >/dev/null
make -C distantspeechrecognition-mirror_b200/csrc -j8 OBJDIR=../build_b TARGET=../libbtkb200_b.so EXTRA="$1" >/dev/null
