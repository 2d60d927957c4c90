# tuning builds of libauv_b200.so (gym_auv_b200/variants/*.so, selected with AUV_B200_LIB=...)
# usage: bash tools/build_variants.sh "NAME:-DFLAG=.. -DFLAG=.." ...
set -e
mkdir -p gym_auv_b200/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-fopenmp -lgomp $flags \
    -o gym_auv_b200/variants/lib_$name.so gym_auv_b200/csrc/auv_kernels.cu &
done
wait
ls -la gym_auv_b200/variants
