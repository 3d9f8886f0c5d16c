set -x
mkdir -p gpurun_out/r2o
timeout 120 python tools/prof_run.py c4 --pairs 256 --reps 2 > gpurun_out/r2o/plain_c4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_wave_linear -s 1 -c 1 -o gpurun_out/r2o/prof_k2_ov_grp python tools/prof_run.py c4 --pairs 256 --reps 2 > gpurun_out/r2o/ncu_ov.log 2>&1
timeout 120 python tools/prof_run.py c3 --pairs 1024 --reps 2 > gpurun_out/r2o/plain_c3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:at_wave_affine -s 1 -c 1 -o gpurun_out/r2o/prof_k2_fj_grp python tools/prof_run.py c3 --pairs 1024 --reps 2 > gpurun_out/r2o/ncu_fj.log 2>&1
