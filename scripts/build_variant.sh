#!/bin/bash
# Kernel experiments: build libfb_<name>.so with extra -D flags for one source file.
#   scripts/build_variant.sh <name> <source.cu> <flags...>    then run with FB_LIB_PATH=fakebob_b200/libfb_<name>.so
set -e
cd "$(dirname "$0")/../fakebob_b200"
name=$1; src=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$src -o /tmp/fbv_$name.o
objs=""
for s in fb_api fb_frontend fb_gmm fb_nes fb_comm fb_ivector fb_enroll; do
  if [ "$s.cu" == "$src" ]; then objs="$objs /tmp/fbv_$name.o"; else objs="$objs csrc/$s.o"; fi
done
nvcc -shared -o libfb_$name.so $objs -lcudart -ldl
echo fakebob_b200/libfb_$name.so
