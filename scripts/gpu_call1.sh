#!/bin/bash
# round-2 GPU call 1: full GPU suite (new fused-objective + gated-gradient tests included), parity of opt-in variants,
# A/B matrix, reference on cuDNN, bench, source-level profiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpus.txt
( timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 ) > gpurun_out/c1_pytest.txt
( OSVOS_TEST_OPTIN=1 timeout 600 python -m pytest tests/test_gpu_optin.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/c1_optin.txt
( timeout 900 bash scripts/ab_matrix.sh ) > gpurun_out/c1_ab_matrix.txt 2>&1
( timeout 300 python scripts/time_reference_gpu.py; timeout 300 python scripts/time_reference_gpu.py --train ) > gpurun_out/c1_ref_gpu.txt 2>&1
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c1_bench.txt 2>gpurun_out/c1_bench.err
( timeout 900 bash scripts/round2_profile.sh ) > gpurun_out/c1_profile.txt 2>&1
tail -5 gpurun_out/c1_pytest.txt; tail -3 gpurun_out/c1_optin.txt; cat gpurun_out/c1_ab_matrix.txt; tail -2 gpurun_out/c1_ref_gpu.txt; tail -c 1500 gpurun_out/c1_bench.err
