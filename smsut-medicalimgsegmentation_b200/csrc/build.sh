#!/bin/bash
# Build libsmsut_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libsmsut_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
OBJS=""
pids=""
for f in conv_tc conv_band wgrad_tc wgrad_band wgrad_hmma conv_direct norm_act resample loss optim det augment coranet; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ common.cuh -nt $f.o ] || [ ../../include/smsut_b200.h -nt $f.o ]; then
    $NVCC $FLAGS "$@" -c $f.cu -o $f.o &
    pids="$pids $!"
  fi
  OBJS="$OBJS $f.o"
done
for p in $pids; do wait $p; done
$NVCC -shared -o $OUT.tmp $OBJS -cudart shared
mv -f $OUT.tmp $OUT      # atomic: a snapshot of the tree never sees a half-written library
echo "built $(readlink -f $OUT)"
