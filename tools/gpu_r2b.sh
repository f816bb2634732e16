#!/bin/bash
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python oracle/gen_ref_golden.py > gpurun_out/${tag}_golden.log 2>&1; echo "golden rc=$?"; tail -30 gpurun_out/${tag}_golden.log
timeout 1500 python -m pytest tests -m gpu -q -k "cpu_conventions or golden or reference_cuda" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${tag}_pytest.log
