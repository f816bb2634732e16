#!/bin/bash
# One GPU visit: parity tests, bench (both arms), ncu launch list, ncu full capture of the evaluation kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2>&1; tail -1 gpurun_out/${tag}_bench_ref.json | cut -c1-300
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/${tag}_bench.json
# ncu passes use 24 pair slots (same kernels, shorter capture); numbers printed under ncu are not bench values
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --pairs 24 --steps 3 --warmup 3 --cpu-budget 0.2 --solves 0 > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hist_sell|k_jac_sell|k_assemble|k_jac_final' -s 8 -c 4 \
  -o gpurun_out/${tag}_prof -f python bench.py --pairs 24 --steps 3 --warmup 3 --cpu-budget 0.2 --solves 0 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
