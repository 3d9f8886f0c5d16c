set -x
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest_gpu.log
for rep in 1 2; do
  timeout 300 python tools/prof_run.py c2 --reps 4 >> gpurun_out/r2f/ab_fma1.log 2>&1
  AT_LIB_PATH=$PWD/aligntools/c_b200/lib_ab0.so timeout 300 python tools/prof_run.py c2 --reps 4 >> gpurun_out/r2f/ab_fma0.log 2>&1
done
timeout 300 python tools/prof_run.py c3 --pairs 1024 --reps 3 >> gpurun_out/r2f/ab_fma1.log 2>&1
AT_LIB_PATH=$PWD/aligntools/c_b200/lib_ab0.so timeout 300 python tools/prof_run.py c3 --pairs 1024 --reps 3 >> gpurun_out/r2f/ab_fma0.log 2>&1
timeout 300 python tools/prof_run.py global --reps 4 >> gpurun_out/r2f/ab_fma1.log 2>&1
AT_LIB_PATH=$PWD/aligntools/c_b200/lib_ab0.so timeout 300 python tools/prof_run.py global --reps 4 >> gpurun_out/r2f/ab_fma0.log 2>&1
