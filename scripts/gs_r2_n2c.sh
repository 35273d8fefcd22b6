mkdir -p gpurun_out
env | grep -i nccl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; tail -c 200 gpurun_out/r2c_bench_n2.json; echo; grep -c "NCCL INFO" gpurun_out/r2c_bench_n2.err; grep "NCCL INFO" gpurun_out/r2c_bench_n2.err | grep -i "nranks" | head -6 | cut -c1-250; ls /tmp/mmc_bench_nccl* 2>/dev/null
