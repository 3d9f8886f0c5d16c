set -x
mkdir -p gpurun_out/r2n4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 --e2e-steps 5 --cfg-steps 2 > gpurun_out/r2n4/bench_n4.json 2> gpurun_out/r2n4/bench_n4.err; echo "rc=$?" >> gpurun_out/r2n4/bench_n4.err
tail -n 3 gpurun_out/r2n4/bench_n4.err
