mkdir -p gpurun_out
for s in -1 25 13 7; do MMC_NUTS_SLICING=$s timeout 300 python scripts/quick_bench.py nuts1 2>&1 | grep nuts_rosen | cut -c1-250; done | tee gpurun_out/r3f_nuts_slicing.log
