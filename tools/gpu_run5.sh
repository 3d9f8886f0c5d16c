set -x
mkdir -p gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest_gpu.log
AT_PIPE_TRACE=1 timeout 300 python bench.py --no-sharded --no-cpu --no-configs --steps 3 --e2e-steps 4 > gpurun_out/r2e/bench_trace.json 2> gpurun_out/r2e/bench_trace.err
timeout 600 python bench.py --no-cpu > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err; echo "bench rc=$?" >> gpurun_out/r2e/bench.err
