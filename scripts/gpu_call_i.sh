#!/bin/bash
# after the clean-up (literal side route, swapped wgrad, side_bwd removed): full GPU suite, benches, launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 ) > gpurun_out/i_pytest.txt
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/i_train480.json 2> gpurun_out/i_train480.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|stage1|fold" -c 120 --csv --log-file gpurun_out/i_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/i_ncu_train.log 2>&1
( timeout 300 python train_online.py --synthetic --no-save 2>&1 | tail -5 ) > gpurun_out/i_online_config3.txt
tail -5 gpurun_out/i_pytest.txt
head -c 330 gpurun_out/i_train480.json | tail -c 200; echo; tail -2 gpurun_out/i_train480.err
grep "side_folded_wgrad\|side_grads_finish" gpurun_out/i_launches_train480.csv | grep gpu__time | cut -d, -f2,12- | cut -c1-40,60- | head
tail -3 gpurun_out/i_online_config3.txt
