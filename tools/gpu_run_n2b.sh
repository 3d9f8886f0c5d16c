set -x
mkdir -p gpurun_out/r2z
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 5 > gpurun_out/r2z/bench_n2.json 2> gpurun_out/r2z/bench_n2.err; echo "rc=$?" >> gpurun_out/r2z/bench_n2.err
tail -3 gpurun_out/r2z/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2z/bench_ref_n2.json 2> gpurun_out/r2z/bench_ref_n2.err; echo "rc=$?" >> gpurun_out/r2z/bench_ref_n2.err
