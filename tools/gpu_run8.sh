set -x
mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/pytest_gpu.log
timeout 300 python tools/prof_run.py c2 --reps 4 > gpurun_out/r2i/plain_c2.log 2>&1
timeout 300 python tools/prof_run.py global --reps 4 > gpurun_out/r2i/plain_global.log 2>&1
timeout 300 python tools/prof_run.py fit --reps 4 > gpurun_out/r2i/plain_fit.log 2>&1
timeout 300 python tools/prof_run.py c3 --pairs 1024 --reps 3 > gpurun_out/r2i/plain_c3.log 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err; echo "bench rc=$?" >> gpurun_out/r2i/bench.err
