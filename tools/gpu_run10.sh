set -x
mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l/pytest_gpu.log
AT_PIPE_TRACE=2 AT_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu --steps 5 > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; echo "bench rc=$?" >> gpurun_out/r2l/bench.err
