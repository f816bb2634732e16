#!/bin/bash
# end-of-round visit: parity tests, both bench arms, ncu launch list + full capture of the evaluation kernels and of kernel 1
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2>&1; tail -1 gpurun_out/${tag}_bench_ref.json | cut -c1-200
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${tag}_bench.json | cut -c1-400
bash tools/gpu_prof.sh ${tag}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_warp_sample_jobs -s 3 -c 1 -f -o gpurun_out/${tag}_k1 python tools/time_k1.py 24 3 > gpurun_out/${tag}_k1_ncu.log 2>&1
timeout 200 python tools/time_k1.py 96 30 > gpurun_out/${tag}_k1_time.json 2>&1; tail -1 gpurun_out/${tag}_k1_time.json
