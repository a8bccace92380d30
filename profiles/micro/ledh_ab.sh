#!/bin/bash
# kernel-only A/B of LEDH flow-kernel builds: ncu durations of k4_ledh_flow_fast at N=2^18 (C5 model, L=8)
#   gpurun -- 'bash profiles/micro/ledh_ab.sh cur 3 4'
for v in "$@"; do
  if [ "$v" = cur ]; then unset SMC_LIB_PATH; else export SMC_LIB_PATH=$PWD/profiles/micro/variants/libsmc_ledh$v.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k4_ledh_flow -c 4 --csv --log-file gpurun_out/ab_$v.csv \
      python bench_configs.py --configs c5 --steps 2 --log2n 18 > /dev/null 2>&1
  echo -n "$v: "; grep k4_ledh gpurun_out/ab_$v.csv | awk -F'","' '{gsub(/[",]/,"",$NF); printf "%s ", $NF} END {print ""}'
done
