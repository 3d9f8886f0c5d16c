set -x
mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest_gpu.log
timeout 300 python tools/prof_run.py global --reps 4 > gpurun_out/r2h/plain_global.log 2>&1
timeout 300 python tools/prof_run.py fit --reps 4 > gpurun_out/r2h/plain_fit.log 2>&1
timeout 300 python tools/prof_run.py global --pairs 1048576 --reps 3 > gpurun_out/r2h/plain_global_1m.log 2>&1
AT_PIPE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 4 --no-configs --no-sharded > gpurun_out/r2h/bench_n2_trace.json 2> gpurun_out/r2h/bench_n2_trace.err
