#!/bin/bash
# round-2 GPU call 9 (1 GPU): stage-1 warps with cheap tap addressing; full suite, A/B, bench, online fine-tune config 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/c9_pytest.txt
export OSVOS_ENV_RELOAD=1
( echo "== OSVOS_FUSE_STAGE1 (1 = default)"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 1 0 || echo FAILED
  for hw in "240 427" "720 1280" "1080 1920"; do echo "== OSVOS_FUSE_STAGE1 at $hw"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 1 0 $hw || echo FAILED; done
) > gpurun_out/c9_ab_matrix.txt 2>&1
unset OSVOS_ENV_RELOAD
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c9_bench.json 2>gpurun_out/c9_bench.err
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/c9_bench_train480.json 2>gpurun_out/c9_bench_train.err
( timeout 200 python scripts/time_forward.py ) > gpurun_out/c9_time_forward.txt 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active
( timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c9_launches_infer480.csv python scripts/one_forward.py ) > gpurun_out/c9_ncu.log 2>&1
( OSVOS_SAVE_ROOT=/tmp/osvos_models timeout 300 python train_online.py --synthetic --no-save ) > gpurun_out/c9_online_finetune_config3.txt 2>&1
( timeout 600 bash scripts/sweep_ncu.sh c9 ) > gpurun_out/c9_sweep.log 2>&1
tail -4 gpurun_out/c9_pytest.txt; cat gpurun_out/c9_ab_matrix.txt; tail -c 300 gpurun_out/c9_bench.err; head -c 400 gpurun_out/c9_bench.json; tail -4 gpurun_out/c9_online_finetune_config3.txt; grep stage1 gpurun_out/c9_launches_infer480.csv | head -3
