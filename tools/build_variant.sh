# usage: bash tools/build_variant.sh NAME -DFLAG=.. -DFLAG2=..   -> variants/lib_NAME.so
N=$1; shift
mkdir -p variants
cd geonomics_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" gnx_api.cu -o ../../variants/lib_$N.so 2>&1 | grep -v "warning #177\|^$\|cur = c->cur\|\^\|Remark" 
