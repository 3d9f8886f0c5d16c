set -x
mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -x -q -k "2bit or k2 or twobit" > gpurun_out/r2v/pytest_2bit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v/pytest_2bit.log
tail -4 gpurun_out/r2v/pytest_2bit.log
timeout 300 python tools/gpu_fuzz.py --seconds 150 --seed 7 > gpurun_out/r2v/fuzz7.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r2v/fuzz7.log
timeout 300 python tools/gpu_fuzz.py --seconds 150 --seed 11 > gpurun_out/r2v/fuzz11.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r2v/fuzz11.log
tail -3 gpurun_out/r2v/fuzz7.log gpurun_out/r2v/fuzz11.log
for rep in 1 2; do
  timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 --twobit >> gpurun_out/r2v/c4_2bit.log 2>&1
  timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2v/c4.log 2>&1
done
grep -H -o '"fill_ms": [0-9.]*\|"gcups": [0-9.]*\|"score_sum": [0-9]*' gpurun_out/r2v/c*.log | paste - - -
