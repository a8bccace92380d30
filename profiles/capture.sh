#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one ncu --set full capture of the SIR step kernels.
#   gpurun --timeout 900 -- 'bash profiles/capture.sh r1d'
# Only after the same bench command has exited 0 without ncu.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_pre.json 2> gpurun_out/${TAG}_bench_pre.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_ncu_launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_|k2_|k3|k_finish' --launch-skip 14 --launch-count 6 \
    -f -o gpurun_out/${TAG}_full python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_ncu_f.log 2>&1
ls -la gpurun_out
