set -x
mkdir -p gpurun_out/r2r
for rep in 1 2; do
for v in default mb5 mb6 mb4u4 mb5u4; do
  if [ $v = default ]; then L=""; else L="$PWD/aligntools/c_b200/lib_$v.so"; fi
  AT_LIB_PATH=$L timeout 200 python tools/prof_run.py c3 --pairs 2048 --reps 3 >> gpurun_out/r2r/c3_$v.log 2>&1
done
done
grep -H gcups gpurun_out/r2r/c3_*.log | cut -c1-260
# where do the warps of the real C3 run wait?  (the earlier captures were one-wave runs of 256-512 pairs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:at_wave_affine -c 1 -o gpurun_out/r2r/k2_c3_2048 -f python tools/prof_run.py c3 --pairs 2048 --reps 1 > gpurun_out/r2r/ncu_c3.log 2>&1
ls -la gpurun_out/r2r
