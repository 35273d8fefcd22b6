mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:dense_gemm_tc_pair -s 2 -c 1 -o gpurun_out/r3_dense_chain python scripts/profile_one.py dense3 2>&1 | tail -2
