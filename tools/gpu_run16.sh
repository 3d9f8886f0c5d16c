set -x
mkdir -p gpurun_out/r2s
for rep in 1 2; do
for o in stripe pair; do
  AT_WAVE_ORDER=$o timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 3 >> gpurun_out/r2s/c3_$o.log 2>&1
  AT_WAVE_ORDER=$o timeout 200 python tools/prof_run.py c3 --pairs 1024 --reps 3 >> gpurun_out/r2s/c3_1024_$o.log 2>&1
  AT_WAVE_ORDER=$o timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 >> gpurun_out/r2s/c4_$o.log 2>&1
  AT_WAVE_ORDER=$o timeout 200 python tools/prof_run.py c5 --pairs 8 --reps 3 >> gpurun_out/r2s/c5_$o.log 2>&1
done
done
grep -H -o '"fill_ms": [0-9.]*\|"gcups": [0-9.]*\|"score_sum": [0-9]*' gpurun_out/r2s/*.log | paste - - -
timeout 1200 python -m pytest tests -m gpu -x -q -k "wave or long or stripe or k2 or full or fuzz or overlap or edit or chunk or fit" > gpurun_out/r2s/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s/pytest_k2.log
tail -5 gpurun_out/r2s/pytest_k2.log
