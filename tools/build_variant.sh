#!/bin/bash
# Builds the library from the current working tree into .tmp_libs/<name>.so (A/B timing on one GPU box):
#   tools/build_variant.sh name ; IMAGEENCODER_B200_LIB=.tmp_libs/name.so python tools/dbg_enc_time.py
set -e
cd "$(dirname "$0")/../imageencoder_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../.tmp_libs
cp ../libimageencoder_b200.so ../../.tmp_libs/$1.so
echo built .tmp_libs/$1.so
