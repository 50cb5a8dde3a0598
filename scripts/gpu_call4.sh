#!/bin/bash
# round-2 GPU call 4: stream-K restricted (tiles >= SMs) with prefetching owner, fused stage 1 v2 (6 stage-1 warps,
# prefetched taps, pipelined TMEM reads), full suite, A/B, bench, resolution sweep with counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu -p no:cacheprovider"
( $T tests/test_gpu_kernels.py -k "stage1 or stream_k" -s 2>&1 | tail -40 ) > gpurun_out/c4_new_kernels.txt
( timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -150 ) > gpurun_out/c4_pytest.txt
export OSVOS_ENV_RELOAD=1
( echo "== OSVOS_STREAMK (1 = default)"; timeout 200 python scripts/ab_env.py OSVOS_STREAMK 1 0 --train || echo FAILED
  echo "== OSVOS_FUSE_STAGE1 (0 = default)"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 0 1 || echo FAILED
  for hw in "240 427" "720 1280" "1080 1920"; do echo "== OSVOS_FUSE_STAGE1 at $hw"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 0 1 $hw || echo FAILED; done
  for hw in "240 427" "720 1280"; do echo "== OSVOS_STREAMK at $hw"; timeout 200 python scripts/ab_env.py OSVOS_STREAMK 1 0 $hw || echo FAILED; done
) > gpurun_out/c4_ab_matrix.txt 2>&1
unset OSVOS_ENV_RELOAD
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c4_bench.json 2>gpurun_out/c4_bench.err
( OSVOS_FUSE_STAGE1=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip dp,gpu_reference,cpu_baseline,e2e_extra ) > gpurun_out/c4_bench_fuse_stage1.json 2>gpurun_out/c4_bench_fuse.err
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/c4_bench_train480.json 2>gpurun_out/c4_bench_train.err
( OSVOS_FUSE_STAGE1=1 timeout 200 python scripts/time_forward.py ) > gpurun_out/c4_time_forward_fuse.txt 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( OSVOS_FUSE_STAGE1=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c4_launches_infer480_fuse.csv python scripts/one_forward.py ) > gpurun_out/c4_ncu_fuse.log 2>&1
( timeout 900 bash scripts/sweep_ncu.sh c4 ) > gpurun_out/c4_sweep.log 2>&1
for f in c4_new_kernels c4_pytest; do echo "== $f"; tail -6 gpurun_out/$f.txt; done
cat gpurun_out/c4_ab_matrix.txt; tail -c 300 gpurun_out/c4_bench.err; cat gpurun_out/c4_sweep_counters.txt | head -80
