set -x
mkdir -p gpurun_out/r2p
for rep in 1 2; do
timeout 120 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2p/c4_default.log 2>&1
AT_LIB_PATH=$PWD/aligntools/c_b200/lib_mb8.so timeout 120 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2p/c4_mb8.log 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q --durations 12 > gpurun_out/r2p/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p/pytest_gpu.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; echo "bench rc=$?" >> gpurun_out/r2p/bench.err
