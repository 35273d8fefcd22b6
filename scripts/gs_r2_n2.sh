mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu_nccl.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 3000 gpurun_out/r2_bench_n2.json; grep -c "NCCL INFO" gpurun_out/r2_bench_n2.err; grep "NCCL INFO.*nranks\|NVLS\|Connected" gpurun_out/r2_bench_n2.err | head -8; tail -3 gpurun_out/r2_bench_n2.err
