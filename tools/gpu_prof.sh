#!/bin/bash
# ncu passes of the evaluation kernels: launch list + full-set capture (24 pair slots; numbers printed under ncu are not bench values)
tag=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --pairs 24 --steps 3 --warmup 3 --cpu-budget 0.2 --solves 0 --c4-pairs 0 --c5 0 --old-gpu 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hist_sell|k_jac_sell|k_assemble|k_jac_final' -s 8 -c 4 \
  -o gpurun_out/${tag}_prof -f $B > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -4
