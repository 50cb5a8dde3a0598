#!/bin/bash
# round-2 GPU call 5 (1 GPU): stream-K removed, fused stage 1 default with 64-byte conv1_1 operand rows (5-stage weight
# ring), batched loads in the tail backward; full suite, A/B, bench (infer + train), sweep with counters, launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu -p no:cacheprovider"
( $T tests/test_gpu_kernels.py -k "stage1" -s 2>&1 | tail -30 ) > gpurun_out/c5_new_kernels.txt
( OSVOS_S1_SW64=0 $T tests/test_gpu_kernels.py -k "stage1" -s 2>&1 | tail -8 ) > gpurun_out/c5_stage1_sw128.txt
( timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -120 ) > gpurun_out/c5_pytest.txt
( timeout 900 bash scripts/ab_matrix.sh ) > gpurun_out/c5_ab_matrix.txt 2>&1
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c5_bench.json 2>gpurun_out/c5_bench.err
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/c5_bench_train480.json 2>gpurun_out/c5_bench_train.err
( timeout 300 python bench.py --steps 20 --warmup 5 --precision fast --skip dp,gpu_reference,cpu_baseline,e2e_extra ) > gpurun_out/c5_bench_fast.json 2>gpurun_out/c5_bench_fast.err
( timeout 200 python scripts/time_forward.py ) > gpurun_out/c5_time_forward.txt 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c5_launches_infer480.csv python scripts/one_forward.py ) > gpurun_out/c5_ncu.log 2>&1
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|sgd|stage1" -c 120 --csv --log-file gpurun_out/c5_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/c5_ncu_train.log 2>&1
( timeout 900 bash scripts/sweep_ncu.sh c5 ) > gpurun_out/c5_sweep.log 2>&1
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:"stage1" -c 1 -f -o gpurun_out/r02e_stage1_fused python scripts/one_forward.py ) > gpurun_out/c5_ncu_full.log 2>&1
for f in c5_new_kernels c5_stage1_sw128 c5_pytest; do echo "== $f"; tail -5 gpurun_out/$f.txt; done
cat gpurun_out/c5_ab_matrix.txt; tail -c 300 gpurun_out/c5_bench.err; head -c 1500 gpurun_out/c5_bench.json
