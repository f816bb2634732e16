#!/bin/bash
# round-2 first visit: regenerate the golden fixtures from the reference's own CUDA code, parity tests, short bench
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 600 python oracle/gen_ref_golden.py > gpurun_out/${tag}_golden.log 2>&1; echo "golden rc=$?"; tail -20 gpurun_out/${tag}_golden.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-budget 2 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/${tag}_bench.json | cut -c1-1500
nproc; free -g | head -2
