set -x
mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests -m gpu -x -q --timeout 240 -k "k2_edge or ragged or full_shape or config3 or config4 or config5 or whitelist or pipelined or edit_unit" > gpurun_out/r2n/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n/pytest_k2.log
timeout 120 python tools/prof_run.py c4 --pairs 256 --reps 3 > gpurun_out/r2n/plain_c4.log 2>&1
timeout 120 python tools/prof_run.py c3 --pairs 2048 --reps 3 > gpurun_out/r2n/plain_c3.log 2>&1
timeout 120 python tools/prof_run.py c3 --pairs 1024 --reps 3 > gpurun_out/r2n/plain_c3_1024.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2n/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n/pytest_gpu.log
