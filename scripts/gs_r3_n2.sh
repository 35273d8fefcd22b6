mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r3_bench_n$N.json 2> gpurun_out/r3_bench_n$N.err; tail -c 300 gpurun_out/r3_bench_n$N.json; echo; grep -c "NCCL INFO" gpurun_out/r3_bench_n$N.err
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_multigpu_nccl.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r3_pytest_gpu_n2.log; fi
