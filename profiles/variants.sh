#!/bin/bash
# profiles/variants.sh NAME "-DICER_DEC_WARPS=4 ..."  -- build libicer_b200 with other tuning constants into
# icer_compression_b200/lib/variants/libicer_b200_NAME.so (select it with ICER_B200_LIB=... for a bench run)
set -e
NAME=$1; DEFS=$2
cd "$(dirname "$0")/../icer_compression_b200/csrc"
mkdir -p build_$NAME ../lib/variants
for f in plan wavelet wavelet_tma coder parse pixfmt api synth; do
  if [ $f = coder ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $DEFS -c $f.cu -o build_$NAME/$f.o
  else
    cp build/$f.o build_$NAME/$f.o
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libicer_b200_$NAME.so build_$NAME/*.o -Xlinker -Bsymbolic
rm -rf build_$NAME
echo built $NAME
