#!/bin/bash
# round-2 GPU call 1: parity of opt-in variants, A/B matrix, reference on cuDNN, baseline bench, source-level profiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpus.txt
( OSVOS_TEST_OPTIN=1 timeout 600 python -m pytest tests/test_gpu_optin.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/c1_optin.txt
( timeout 900 bash scripts/ab_matrix.sh ) > gpurun_out/c1_ab_matrix.txt 2>&1
( timeout 300 python scripts/time_reference_gpu.py; timeout 300 python scripts/time_reference_gpu.py --train ) > gpurun_out/c1_ref_gpu.txt 2>&1
( timeout 300 python bench.py --steps 200 --warmup 20 ) > gpurun_out/c1_bench.txt 2>gpurun_out/c1_bench.err
( timeout 900 bash scripts/round2_profile.sh ) > gpurun_out/c1_profile.txt 2>&1
tail -3 gpurun_out/c1_optin.txt; cat gpurun_out/c1_ab_matrix.txt; cat gpurun_out/c1_ref_gpu.txt | tail -2
