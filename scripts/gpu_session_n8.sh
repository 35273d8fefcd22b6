set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 1500 gpurun_out/bench_n8.json
timeout 900 $TR scripts/bench_configs.py --no-cpu > gpurun_out/all_configs_n8.jsonl 2> gpurun_out/all_configs_n8.err; cut -c1-600 gpurun_out/all_configs_n8.jsonl; tail -5 gpurun_out/all_configs_n8.err
