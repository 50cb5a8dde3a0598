#!/bin/bash
# folded side-branch backward v2 + tap-row wgrad of conv1_2: parity tests, fwd+bwd A/B, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_side_folded.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/d_pytest_new.txt
( timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_objective.py tests/test_gpu_optim.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/d_pytest_bwd.txt
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/d_train480_folded.json 2> gpurun_out/d_train480_folded.err
( OSVOS_WGRAD_ROWS=0 timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/d_train480_folded_pairs.json 2>/dev/null
( OSVOS_SIDE_BWD=literal OSVOS_WGRAD_ROWS=0 timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/d_train480_literal_pairs.json 2>/dev/null
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|stage1|fold" -c 120 --csv --log-file gpurun_out/d_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/d_ncu_train.log 2>&1
tail -4 gpurun_out/d_pytest_new.txt; tail -4 gpurun_out/d_pytest_bwd.txt
for f in d_train480_folded d_train480_folded_pairs d_train480_literal_pairs; do echo $f; head -c 330 gpurun_out/$f.json | tail -c 200; echo; done
tail -3 gpurun_out/d_train480_folded.err
