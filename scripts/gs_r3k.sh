mkdir -p gpurun_out
MMC_TC_VERBOSE=1 timeout 120 python scripts/quick_bench.py dense_tc 2>&1 | grep '"path": 3\|minimcmc' | cut -c1-330 | tee gpurun_out/r3k_dense_quad.log
MMC_TC_QUAD=1 MMC_TC_VERBOSE=1 timeout 120 python scripts/quick_bench.py dense_tc 2>&1 | grep '"path": 3\|minimcmc' | cut -c1-330 | tee -a gpurun_out/r3k_dense_quad.log
MMC_TC_QUAD=0 timeout 120 python scripts/quick_bench.py dense_tc 2>&1 | grep '"path": 3' | cut -c1-330 | tee -a gpurun_out/r3k_dense_quad.log
