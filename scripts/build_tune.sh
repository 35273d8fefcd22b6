#!/bin/bash
# Tuning builds of libminimcmc.so: recompiles one translation unit with extra -D flags and links it with the default
# objects.  usage: scripts/build_tune.sh <name> <unit.cu> <flags...>  ->  tune/libminimcmc_<name>.so (MMC_LIB_PATH)
set -e
cd "$(dirname "$0")/../mini_mcmc_b200/csrc"
name=$1; unit=$2; shift 2
mkdir -p ../../tune ../../build/tune
obj=../../build/tune/${name}_${unit%.cu}.o
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fvisibility=default,-ffp-contract=off --expt-relaxed-constexpr -Xptxas -v "$@" -c $unit -o $obj 2> ../../build/tune/${name}.ptxas.log
objs=$(ls ../../build/obj/*.o | grep -v "/${unit%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o ../../tune/libminimcmc_${name}.so $objs $obj -lcuda -lpthread
echo built tune/libminimcmc_${name}.so
