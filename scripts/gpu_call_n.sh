#!/bin/bash
# bounded buffer-bound graphs: forward tests; BASELINE configs[4] sweep with counters on the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_output.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/n_pytest.txt
tail -3 gpurun_out/n_pytest.txt
( timeout 300 python bench.py --steps 20 --warmup 5 --skip cpu_baseline,gpu_reference,dp,parity ) > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
head -c 300 gpurun_out/n_bench.json | tail -c 160; echo; tail -2 gpurun_out/n_bench.err
bash scripts/sweep_ncu.sh r02n
