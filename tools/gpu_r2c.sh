#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python tools/time_setup.py 256 2>&1 | tee gpurun_out/${tag}_setup.log | tail
timeout 600 python tools/time_setup.py 128 16 10 2>&1 | tee -a gpurun_out/${tag}_setup.log | tail -5
