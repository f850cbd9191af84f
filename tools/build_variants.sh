#!/bin/sh
# development helper: builds tuning variants of the library into build_variants/ (git-ignored, they travel to the GPU box).
#   sh tools/build_variants.sh name1:"-DA=1 -DB=2" name2:"..."
set -e
mkdir -p build_variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( ACLB200_OUT=$PWD/build_variants/lib_$name.so sh acl_b200/csrc/build.sh $flags 2>&1 | grep -E "error|pipeline_kernelILi1ELb0ELb0ELb0E|built" | grep -v Compiling | head -3
    ACLB200_OUT=$PWD/build_variants/lib_$name.so sh acl_b200/csrc/build.sh $flags 2>&1 | grep -A2 "Function properties.*pipeline_kernelILi1ELb0ELb0ELb0E" | tail -2 ) &
done
wait
