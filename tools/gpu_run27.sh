set -x
mkdir -p gpurun_out/r2pf
for rep in 1 2; do
for v in default pf8 pf8b pf16; do
  if [ $v = default ]; then L=""; else L="$PWD/aligntools/c_b200/lib_$v.so"; fi
  AT_LIB_PATH=$L timeout 200 python tools/prof_run.py c4 --pairs 256 --reps 3 --twobit >> gpurun_out/r2pf/c4_$v.log 2>&1
  AT_LIB_PATH=$L timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 2 --twobit >> gpurun_out/r2pf/c3_$v.log 2>&1
done
done
grep -H -o '"traceback_ms": [0-9.]*\|"gcups": [0-9.]*\|"score_sum": [0-9]*' gpurun_out/r2pf/c*.log | paste - - -
