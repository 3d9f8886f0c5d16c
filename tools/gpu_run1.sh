set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a/gpu.txt; nproc >> gpurun_out/r2a/gpu.txt; free -g >> gpurun_out/r2a/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?" >> gpurun_out/r2a/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a/bench_ref.json 2>> gpurun_out/r2a/bench.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_peak tools/int_peak.cu && /tmp/int_peak > gpurun_out/r2a/int_peak.txt 2>&1
timeout 300 python tools/prof_run.py c2 --reps 2 > gpurun_out/r2a/plain_c2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_fill_affine -s 1 -c 1 -o gpurun_out/r2a/prof_k1_tag python tools/prof_run.py c2 --reps 2 > gpurun_out/r2a/ncu_k1.log 2>&1
timeout 300 python tools/prof_run.py c3 --pairs 256 --reps 2 > gpurun_out/r2a/plain_c3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_wave_affine -s 1 -c 1 -o gpurun_out/r2a/prof_k2_fj python tools/prof_run.py c3 --pairs 256 --reps 2 > gpurun_out/r2a/ncu_k2.log 2>&1
ls -la gpurun_out/r2a
