#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q3_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q3_pytest.log
timeout 300 python tools/time_config.py 480 640 4 16 96 20 2>&1 | grep 'path=sorted'
[ -f build/libvar_asmtrace.so ] && NID_B200_LIB=$PWD/build/libvar_asmtrace.so python tools/asm_trace.py 2>&1 | tail -16 | sort -k5 | cut -c1-150 | head -6
bash tools/gpu_lat.sh task_px=16 task_px=16,asm_wide=0
