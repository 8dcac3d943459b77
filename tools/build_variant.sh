#!/bin/bash
# tools/build_variant.sh NAME "-DB200_G1_MINB=5 ..."  ->  gpurun_variants/libb200kzg_NAME.so  (tuning builds, not the product)
set -e
name=$1; shift
out=go_kzg_b200/lib/variants; mkdir -p $out/obj_$name
for u in api kernels_fr kernels_g1; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -I include $@ \
     -c go_kzg_b200/csrc/$u.cu -o $out/obj_$name/$u.o 2>&1 | grep -v deprecated &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out/libb200kzg_$name.so $out/obj_$name/*.o -lcudart
echo built $out/libb200kzg_$name.so
