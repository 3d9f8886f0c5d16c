set -x
mkdir -p gpurun_out/r2w
for lag in 0 256 1024; do
  AT_WAVE_START_LAG=$lag timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2w/c4_lag$lag.log 2>&1
  AT_WAVE_START_LAG=$lag timeout 200 python tools/prof_run.py c3 --pairs 1024 --reps 3 >> gpurun_out/r2w/c3_1024_lag$lag.log 2>&1
  AT_WAVE_START_LAG=$lag timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 3 >> gpurun_out/r2w/c3_2048_lag$lag.log 2>&1
done
grep -H -o '"fill_ms": [0-9.]*\|"gcups": [0-9.]*\|"score_sum": [0-9]*' gpurun_out/r2w/c*.log | paste - - -
