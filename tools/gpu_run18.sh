set -x
mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests -m gpu -x -q -k "2bit or k2 or twobit or whitelist or fuzz or config3 or config4 or edit" > gpurun_out/r2u/pytest_2bit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u/pytest_2bit.log
tail -15 gpurun_out/r2u/pytest_2bit.log
timeout 200 python tools/gpu_fuzz.py --seconds 100 --seed 7 > gpurun_out/r2u/fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r2u/fuzz.log
tail -4 gpurun_out/r2u/fuzz.log
for rep in 1 2; do
  timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 3 >> gpurun_out/r2u/c3.log 2>&1
  timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 3 --twobit >> gpurun_out/r2u/c3_2bit.log 2>&1
  timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2u/c4.log 2>&1
  timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 --twobit >> gpurun_out/r2u/c4_2bit.log 2>&1
  timeout 200 python tools/prof_run.py c5 --pairs 8 --reps 3 --twobit >> gpurun_out/r2u/c5_2bit.log 2>&1
done
grep -H -o '"fill_ms": [0-9.]*\|"gcups": [0-9.]*\|"score_sum": [0-9]*' gpurun_out/r2u/c*.log | paste - - -
