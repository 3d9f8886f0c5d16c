set -x
mkdir -p gpurun_out/r2h
timeout 900 python tools/parity_full.py --twobit --max-seconds 540 --c3 8 --c4 6 --c5 2 --out gpurun_out/r2h/parity_full.json > gpurun_out/r2h/parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2h/parity.log
tail -n 4 gpurun_out/r2h/parity.log
timeout 400 python tools/gpu_fuzz.py --seconds 240 --seed 23 > gpurun_out/r2h/fuzz23.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r2h/fuzz23.log
tail -n 3 gpurun_out/r2h/fuzz23.log
