set -x
mkdir -p gpurun_out/r2j
nvidia-smi topo -m > gpurun_out/r2j/topo.txt 2>&1; nproc >> gpurun_out/r2j/topo.txt; free -g >> gpurun_out/r2j/topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j/bench_n8.json 2> gpurun_out/r2j/bench_n8.err; echo "rc=$?" >> gpurun_out/r2j/bench_n8.err
AT_PIPE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 3 --no-configs --no-sharded > gpurun_out/r2j/bench_n8_trace.json 2> gpurun_out/r2j/bench_n8_trace.err
