#!/bin/bash
# build a kernel variant into _var/<name>/libkpp_gpu.so :  tools/build_variant.sh name [ENV=1 ...]
set -e
name=$1; shift
mkdir -p _var/$name
env "$@" python -m mckpp_f90_b200.build --force > /dev/null
cp mckpp_f90_b200/libkpp_gpu.so _var/$name/libkpp_gpu.so
