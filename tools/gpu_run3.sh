set -x
mkdir -p gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest_gpu.log
timeout 600 python bench.py --no-sharded --no-cpu > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err; echo "bench rc=$?" >> gpurun_out/r2c/bench.err
timeout 300 python tools/prof_run.py c2 --reps 2 > gpurun_out/r2c/plain_c2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_fill_affine -s 1 -c 1 -o gpurun_out/r2c/prof_k1_cell python tools/prof_run.py c2 --reps 2 > gpurun_out/r2c/ncu_k1.log 2>&1
timeout 300 python tools/prof_run.py c3 --pairs 256 --reps 2 > gpurun_out/r2c/plain_c3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_wave_affine -s 1 -c 1 -o gpurun_out/r2c/prof_k2_cell python tools/prof_run.py c3 --pairs 256 --reps 2 > gpurun_out/r2c/ncu_k2.log 2>&1
