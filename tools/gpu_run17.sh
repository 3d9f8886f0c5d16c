set -x
mkdir -p gpurun_out/r2t
AT_PIPE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 3 > gpurun_out/r2t/bench_trace.json 2> gpurun_out/r2t/trace.err
tail -c 3000 gpurun_out/r2t/trace.err
