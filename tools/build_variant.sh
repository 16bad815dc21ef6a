#!/bin/bash
# Builds scratch/lib_<name>.so: the library with gicp.cu compiled under extra -D flags (A/B runs through ORPCD_B200_LIB).
# usage: tools/build_variant.sh name [-DX=1 ...] [-- file.cu]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/multi-scale-pointcloud-registration_b200/csrc
name=$1; shift
FILE=${VARIANT_FILE:-gicp}
mkdir -p $ROOT/scratch
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ --cudart shared"
/usr/local/cuda/bin/nvcc $FLAGS "$@" -c -o $ROOT/scratch/${FILE}_$name.o $CS/$FILE.cu
OBJS=""
for f in knn covariance gicp fpfh fgr downsample peak; do
  if [ $f = $FILE ]; then OBJS="$OBJS $ROOT/scratch/${FILE}_$name.o"; else OBJS="$OBJS $CS/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart shared -ccbin /usr/bin/g++ -o $ROOT/scratch/lib_$name.so $OBJS
/usr/local/cuda/bin/cuobjdump -res-usage $ROOT/scratch/${FILE}_$name.o | grep -A1 "${VARIANT_KERNEL:-gicp_persistent_kernelILi1}" | tail -1
