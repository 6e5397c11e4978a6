#!/usr/bin/env bash
# Builds a tuning variant of the library next to the product build: tools/build_variant.sh <name> <nvcc -D flags...>
# -> jvector_b200/lib_<name>/libjvector_b200.so, loaded with JV_B200_SO=<that path> (tools/profile_search.py --envs ...)
cd "$(dirname "$0")/.."
name=$1; shift
JV_B200_LIBDIR=$PWD/jvector_b200/lib_$name JV_NVCC_EXTRA="$*" python -m jvector_b200.build
