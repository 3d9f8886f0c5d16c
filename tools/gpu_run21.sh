set -x
mkdir -p gpurun_out/r2x
for rep in 1 2; do
  timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2x/c4.log 2>&1
  timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 3 >> gpurun_out/r2x/c3.log 2>&1
done
grep -H -o '"fill_ms": [0-9.]*\|"traceback_ms": [0-9.]*\|"gcups": [0-9.]*' gpurun_out/r2x/c*.log | paste - - -
timeout 1500 python -m pytest tests -m gpu -x -q --durations 8 > gpurun_out/r2x/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x/pytest_gpu.log
tail -14 gpurun_out/r2x/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2x/bench.json 2> gpurun_out/r2x/bench.err; echo "bench rc=$?" >> gpurun_out/r2x/bench.err
tail -3 gpurun_out/r2x/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x/bench_ref.json 2> gpurun_out/r2x/bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2x/bench_ref.err
