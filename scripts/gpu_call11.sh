#!/bin/bash
# round-2 GPU call 11 (1 GPU): fused stage 1 v4 (im2col rows built by the conv1_2 epilogue warps, three tiles ahead)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu -p no:cacheprovider"
( $T tests/test_gpu_kernels.py -k "stage1" -s 2>&1 | tail -20 ) > gpurun_out/c11_stage1.txt
( timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/c11_pytest.txt
export OSVOS_ENV_RELOAD=1
( echo "== OSVOS_FUSE_STAGE1 (1 = default)"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 1 0 || echo FAILED
  for hw in "240 427" "720 1280" "1080 1920"; do echo "== OSVOS_FUSE_STAGE1 at $hw"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 1 0 $hw || echo FAILED; done
) > gpurun_out/c11_ab_matrix.txt 2>&1
unset OSVOS_ENV_RELOAD
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c11_bench.json 2>gpurun_out/c11_bench.err
( timeout 200 python scripts/time_forward.py ) > gpurun_out/c11_time_forward.txt 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active
( timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c11_launches_infer480.csv python scripts/one_forward.py ) > gpurun_out/c11_ncu.log 2>&1
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:"stage1|side_conv" -c 2 -f -o gpurun_out/r02j_stage1_side python scripts/one_forward.py ) > gpurun_out/c11_ncu_full.log 2>&1
tail -3 gpurun_out/c11_stage1.txt; tail -4 gpurun_out/c11_pytest.txt; cat gpurun_out/c11_ab_matrix.txt; tail -c 300 gpurun_out/c11_bench.err; head -c 400 gpurun_out/c11_bench.json; grep stage1 gpurun_out/c11_launches_infer480.csv | grep "gpu__time" | head -2
