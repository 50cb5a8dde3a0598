#!/bin/bash
# merged side-conv launch (inference + training forward), finish kernel v2: parity, A/B, profile of side_folded_wgrad
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py tests/test_gpu_side_folded.py tests/test_gpu_objective.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/e_pytest.txt
( timeout 300 python scripts/ab_env.py OSVOS_SIDE_MULTI 1 0 ) > gpurun_out/e_ab_side_multi_480.txt 2>&1
( timeout 300 python scripts/ab_env.py OSVOS_SIDE_MULTI 1 0 240 427 ) > gpurun_out/e_ab_side_multi_240.txt 2>&1
( timeout 300 python bench.py --steps 20 --warmup 5 --skip cpu_baseline,gpu_reference,dp,e2e_extra ) > gpurun_out/e_bench_infer.json 2> gpurun_out/e_bench_infer.err
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/e_train480.json 2> gpurun_out/e_train480.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1|fold" -c 40 --csv --log-file gpurun_out/e_launches_infer480.csv python scripts/one_forward.py ) > gpurun_out/e_ncu_infer.log 2>&1
( timeout 400 ncu --set full --import-source on --clock-control none -k regex:"side_folded_wgrad" -c 2 -o gpurun_out/e_side_wgrad python scripts/one_train_step.py ) > gpurun_out/e_ncu_wgrad.log 2>&1
tail -4 gpurun_out/e_pytest.txt; cat gpurun_out/e_ab_side_multi_480.txt gpurun_out/e_ab_side_multi_240.txt | grep fps
head -c 260 gpurun_out/e_bench_infer.json | tail -c 160; echo; head -c 260 gpurun_out/e_train480.json | tail -c 160; echo
ls -la gpurun_out/e_side_wgrad.ncu-rep
