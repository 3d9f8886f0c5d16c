set -x
mkdir -p gpurun_out/r2d
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest_gpu.log
timeout 300 python tools/prof_run.py c4 --pairs 256 --reps 3 > gpurun_out/r2d/plain_c4.log 2>&1
timeout 300 python tools/prof_run.py c3 --pairs 1024 --reps 3 > gpurun_out/r2d/plain_c3.log 2>&1
timeout 300 python tools/prof_run.py c3 --pairs 2048 --reps 3 > gpurun_out/r2d/plain_c3_2048.log 2>&1
AT_PIPE_TRACE=1 timeout 300 python bench.py --no-sharded --no-cpu --no-configs --steps 3 --e2e-steps 3 > gpurun_out/r2d/bench_trace.json 2> gpurun_out/r2d/bench_trace.err
timeout 300 python tools/prof_run.py c4 --pairs 256 --reps 2 > gpurun_out/r2d/plain_c4b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_wave_linear -s 1 -c 1 -o gpurun_out/r2d/prof_k2_ov python tools/prof_run.py c4 --pairs 256 --reps 2 > gpurun_out/r2d/ncu_ov.log 2>&1
