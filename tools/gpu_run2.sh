set -x
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest_gpu.log
timeout 600 python bench.py --no-sharded > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; echo "bench rc=$?" >> gpurun_out/r2b/bench.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_peak tools/int_peak.cu && /tmp/int_peak > gpurun_out/r2b/int_peak.txt 2>&1
timeout 300 python tools/prof_run.py global --reps 3 > gpurun_out/r2b/plain_global.log 2>&1
timeout 300 python tools/prof_run.py fit --reps 3 > gpurun_out/r2b/plain_fit.log 2>&1
