# usage: bash scripts/gs_nuts_tune.sh <tag> <lib names...>   (default lib first)
tag=$1; shift
mkdir -p gpurun_out
( timeout 300 python scripts/quick_bench.py nuts1 2>&1 | cut -c1-330
for n in "$@"; do MMC_LIB_PATH=$PWD/tune/libminimcmc_$n.so timeout 300 python scripts/quick_bench.py nuts1 2>&1 | cut -c1-330; done ) | tee gpurun_out/nuts_tune_$tag.log
