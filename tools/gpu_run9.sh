set -x
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k/pytest_gpu.log
AT_PIPE_TRACE=2 timeout 300 python bench.py --no-sharded --no-cpu --no-configs --steps 3 --e2e-steps 4 > gpurun_out/r2k/bench_trace.json 2> gpurun_out/r2k/bench_trace.err
AT_SYNC=block timeout 300 python bench.py --no-sharded --no-cpu --no-configs --steps 3 --e2e-steps 6 > gpurun_out/r2k/bench_block.json 2> gpurun_out/r2k/bench_block.err
timeout 600 python bench.py --no-cpu > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err; echo "bench rc=$?" >> gpurun_out/r2k/bench.err
