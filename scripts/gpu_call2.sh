#!/bin/bash
# round-2 GPU call 2: new kernels (stream-K, folded side branch, fused objective, two-phase tail) - isolated test
# processes first, then the full suite, A/B matrix, bench and launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu -p no:cacheprovider"
( $T tests/test_gpu_kernels.py -k "stream_k or folded or tail" -s 2>&1 | tail -25 ) > gpurun_out/c2_new_kernels.txt
( $T tests/test_gpu_objective.py -s 2>&1 | tail -30 ) > gpurun_out/c2_objective.txt
( $T tests/test_gpu_backward.py -k "gates" -s 2>&1 | tail -25 ) > gpurun_out/c2_gates.txt
( timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -120 ) > gpurun_out/c2_pytest.txt
( timeout 900 bash scripts/ab_matrix.sh ) > gpurun_out/c2_ab_matrix.txt 2>&1
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c2_bench.json 2>gpurun_out/c2_bench.err
( OSVOS_STREAMK=0 OSVOS_FOLD_SIDE=0 timeout 300 python bench.py --steps 20 --warmup 5 --skip dp,parity,gpu_reference,cpu_baseline,e2e_extra ) > gpurun_out/c2_bench_nostreamk_nofold.json 2>/dev/null
( timeout 200 python scripts/time_forward.py ) > gpurun_out/c2_time_forward.txt 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size --clock-control none -c 40 --csv --log-file gpurun_out/c2_launches_infer480.csv python scripts/one_forward.py ) > gpurun_out/c2_ncu.log 2>&1
( timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size --clock-control none -c 200 --csv --log-file gpurun_out/c2_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/c2_ncu_train.log 2>&1
for f in c2_new_kernels c2_objective c2_gates c2_pytest; do echo "== $f"; tail -4 gpurun_out/$f.txt; done
cat gpurun_out/c2_ab_matrix.txt; tail -c 600 gpurun_out/c2_bench.err
