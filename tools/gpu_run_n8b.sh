set -x
mkdir -p gpurun_out/r2m
AT_PIPE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2m/bench_n8.json 2> gpurun_out/r2m/bench_n8.err; echo "rc=$?" >> gpurun_out/r2m/bench_n8.err
AT_SYNC=spin timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 6 --no-configs --no-sharded > gpurun_out/r2m/bench_n8_spin.json 2> gpurun_out/r2m/bench_n8_spin.err
