mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 500 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r3_bench_n8.json 2> gpurun_out/r3_bench_n8.err; tail -c 600 gpurun_out/r3_bench_n8.json; echo; grep -c "NCCL INFO" gpurun_out/r3_bench_n8.err
