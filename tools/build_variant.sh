#!/bin/bash
# tools/build_variant.sh NAME "NVCC FLAGS": build the library with extra flags into _ab/libNAME.so
# (A/B timing on the GPU box: tools/ab_step.sh NAME ..., or PYPDE_B200_LIB=$PWD/_ab/libNAME.so python ...).
# The default build is restored afterwards.
set -e
cd "$(dirname "$0")/.."
mkdir -p _ab
PDE_NVCC_EXTRA="$2" python -m pypde_b200.build --force > /dev/null
cp pypde_b200/_lib/libpypde_b200.so _ab/lib$1.so
python -m pypde_b200.build --force > /dev/null
echo built _ab/lib$1.so with "$2" and restored the default build
