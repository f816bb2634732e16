#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${tag}_pytest.log
timeout 300 python tools/time_setup.py 256 2>&1 | tail -2
timeout 300 python tools/time_single.py 2>&1 | tail -6
NID_OPTS=lm_reuse=0 timeout 300 python tools/time_single.py 2>&1 | tail -1
