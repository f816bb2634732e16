#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_config.py 480 640 4 16 96 20 2>&1 | grep 'path=sorted'
[ -f build/libvar_asmtrace.so ] && NID_B200_LIB=$PWD/build/libvar_asmtrace.so python tools/asm_trace.py 2>&1 | tail -16 | sort -k5 | cut -c1-130
NID_LM_TIMES=1 timeout 300 python tools/time_single.py 480 640 4 16 2>&1 | tail -4 | tr '\n' ' '; echo
