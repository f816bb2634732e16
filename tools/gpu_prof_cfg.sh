#!/bin/bash
# ncu --set full capture of the pixel kernels for one geometry: tools/gpu_prof_cfg.sh <tag> rows cols cell bins
tag=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hist_sell|k_jac_sell|k_assemble_warp' -s 6 -c 3 \
  -o gpurun_out/${tag}_prof -f python tools/time_config.py "$@" 24 2 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${tag}
