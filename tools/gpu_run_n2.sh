set -x
mkdir -p gpurun_out/r2g
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g/bench_n2.json 2> gpurun_out/r2g/bench_n2.err; echo "rc=$?" >> gpurun_out/r2g/bench_n2.err
nvidia-smi topo -m > gpurun_out/r2g/topo.txt 2>&1; nproc >> gpurun_out/r2g/topo.txt; numactl -H >> gpurun_out/r2g/topo.txt 2>&1
