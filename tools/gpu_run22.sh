set -x
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests -m gpu -x -q -k "pipe or shot or slice or chunk" > gpurun_out/r2y/pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y/pytest_pipe.log
tail -3 gpurun_out/r2y/pytest_pipe.log
AT_PIPE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 5 > gpurun_out/r2y/bench_trace.json 2> gpurun_out/r2y/trace.err
AT_PIPE_NO_RAMP_DOWN=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 5 > gpurun_out/r2y/bench_noramp.json 2> gpurun_out/r2y/noramp.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 5 > gpurun_out/r2y/bench_ramp.json 2> gpurun_out/r2y/ramp.err
AT_PIPE_NO_RAMP_DOWN=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 5 > gpurun_out/r2y/bench_noramp2.json 2> gpurun_out/r2y/noramp2.err
