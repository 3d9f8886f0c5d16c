set -x
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -x -q --durations 5 > gpurun_out/r2f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest_gpu.log
tail -10 gpurun_out/r2f/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; echo "bench rc=$?" >> gpurun_out/r2f/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f/bench_ref.json 2> gpurun_out/r2f/bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2f/bench_ref.err
# launch list of the bench command (the unprofiled run of the same command first)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 2 > gpurun_out/r2f/bench_short.json 2> gpurun_out/r2f/bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-sharded --e2e-steps 2 > gpurun_out/r2f/ncu_bench.log 2>&1
# full captures of the shipped kernels (the commands have exited 0 unprofiled above / in earlier calls)
timeout 200 python tools/prof_run.py c2 --twobit --reps 1 > gpurun_out/r2f/prof_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_fill_affine -c 1 -o gpurun_out/r2f/k1_c2 -f python tools/prof_run.py c2 --twobit --reps 1 > gpurun_out/r2f/ncu_c2.log 2>&1
timeout 200 python tools/prof_run.py c3 --pairs 2048 --twobit --reps 1 > gpurun_out/r2f/prof_c3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_wave_affine -c 1 -o gpurun_out/r2f/k2_c3 -f python tools/prof_run.py c3 --pairs 2048 --twobit --reps 1 > gpurun_out/r2f/ncu_c3.log 2>&1
timeout 200 python tools/prof_run.py c4 --pairs 256 --twobit --reps 1 > gpurun_out/r2f/prof_c4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_wave_linear -c 1 -o gpurun_out/r2f/k2_c4 -f python tools/prof_run.py c4 --pairs 256 --twobit --reps 1 > gpurun_out/r2f/ncu_c4.log 2>&1
for k in k1_c2 k2_c3 k2_c4; do
  [ -f gpurun_out/r2f/$k.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r2f/$k.ncu-rep gpurun_out/r2f/$k.csv > /dev/null 2>> gpurun_out/r2f/summary.err
done
rm -f gpurun_out/r2f/k1_c2.ncu-rep gpurun_out/r2f/k2_c4.ncu-rep
ls -la gpurun_out/r2f; du -sh gpurun_out
