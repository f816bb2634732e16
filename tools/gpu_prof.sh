#!/bin/bash
# ncu full capture of kernels matching $2 (regex) during a short bench; report -> gpurun_out/$1.ncu-rep
tag=$1; pat=$2; shift; shift
mkdir -p gpurun_out
NID_OPTS=$NID_OPTS timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s ${SKIP:-8} -c ${COUNT:-4} \
  -o gpurun_out/${tag} -f python bench.py --steps 3 --warmup 3 --cpu-budget 0.2 "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log | cut -c1-300
ls -la gpurun_out/${tag}.ncu-rep
