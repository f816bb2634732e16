#!/bin/bash
# full GPU parity tests, then the C2 throughput split, the assembly trace and the single-pair latencies
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q2_pytest.log
timeout 300 python tools/time_config.py 480 640 4 16 96 20 2>&1 | grep 'path=sorted'
timeout 300 python tools/time_config.py 480 640 16 10 96 20 2>&1 | grep 'path=sorted want_jac=1'
[ -f build/libvar_asmtrace.so ] && NID_B200_LIB=$PWD/build/libvar_asmtrace.so python tools/asm_trace.py 2>&1 | tail -16 | sort -k5 | awk '{print $3, $9, $10, $11, $12, $13, $14}' | tr '\n' ';'; echo
NID_LM_TIMES=1 timeout 300 python tools/time_single.py 480 640 4 16 2>&1 | tail -4 | tr '\n' ' '; echo
