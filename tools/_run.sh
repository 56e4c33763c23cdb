python -m pytest tests -x -q -m gpu 2>&1 | tail -3
DFPSR_ASYNC=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 1500 gpurun_out/bench_r2_n2.err; head -c 3000 gpurun_out/bench_r2_n2.json
