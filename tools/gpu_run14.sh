set -x
mkdir -p gpurun_out/r2q
timeout 330 python tools/gpu_fuzz.py --seconds 240 > gpurun_out/r2q/fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r2q/fuzz.log
timeout 900 python tools/parity_full.py --twobit --max-seconds 540 --out gpurun_out/r2q/parity_full.json > gpurun_out/r2q/parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2q/parity.log
# launch list of the bench command (unprofiled run first)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 2 > gpurun_out/r2q/bench_short.json 2> gpurun_out/r2q/bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 2 > gpurun_out/r2q/ncu_bench.log 2>&1
# full captures of the shipped kernels (each command has exited 0 unprofiled in earlier calls and in the test run)
timeout 200 python tools/prof_run.py c2 --twobit --reps 1 > gpurun_out/r2q/prof_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_fill_affine -c 1 -o gpurun_out/r2q/k1_c2 -f python tools/prof_run.py c2 --twobit --reps 1 > gpurun_out/r2q/ncu_c2.log 2>&1
timeout 200 python tools/prof_run.py c3 --pairs 512 --reps 1 > gpurun_out/r2q/prof_c3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_wave_affine -c 1 -o gpurun_out/r2q/k2_c3 -f python tools/prof_run.py c3 --pairs 512 --reps 1 > gpurun_out/r2q/ncu_c3.log 2>&1
timeout 200 python tools/prof_run.py c4 --pairs 128 --reps 1 > gpurun_out/r2q/prof_c4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_wave_linear -c 1 -o gpurun_out/r2q/k2_c4 -f python tools/prof_run.py c4 --pairs 128 --reps 1 > gpurun_out/r2q/ncu_c4.log 2>&1
for k in k1_c2 k2_c3 k2_c4; do
  [ -f gpurun_out/r2q/$k.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r2q/$k.ncu-rep gpurun_out/r2q/$k.csv > /dev/null 2>> gpurun_out/r2q/summary.err
  rm -f gpurun_out/r2q/$k.ncu-rep
done
ls -la gpurun_out/r2q
du -sh gpurun_out
