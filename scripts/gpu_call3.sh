#!/bin/bash
# round-2 GPU call 3: call 2 again after the stream-K workspace fix + fused stage-1 kernel, deeper folded-side ring,
# transposing bias-gradient reduction
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu -p no:cacheprovider"
( $T tests/test_gpu_kernels.py -k "stage1 or stream_k or folded" -s 2>&1 | tail -40 ) > gpurun_out/c3_new_kernels.txt
( $T tests/test_gpu_backward.py -k "gates" -s 2>&1 | grep -v Warning | head -150 ) > gpurun_out/c3_gates.txt
( timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -150 ) > gpurun_out/c3_pytest.txt
export OSVOS_ENV_RELOAD=1
( for sw in OSVOS_STREAMK OSVOS_FOLD_SIDE OSVOS_HALO_LEAN; do echo "== $sw (1 = default)"; timeout 200 python scripts/ab_env.py $sw 1 0 --train || echo FAILED; done
  echo "== OSVOS_FUSE_STAGE1 (0 = default)"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 0 1 || echo FAILED
  for hw in "240 427" "720 1280"; do echo "== OSVOS_STREAMK at $hw"; timeout 200 python scripts/ab_env.py OSVOS_STREAMK 1 0 $hw || echo FAILED; done
  for hw in "240 427" "720 1280"; do echo "== OSVOS_FUSE_STAGE1 at $hw"; timeout 200 python scripts/ab_env.py OSVOS_FUSE_STAGE1 0 1 $hw || echo FAILED; done
) > gpurun_out/c3_ab_matrix.txt 2>&1
unset OSVOS_ENV_RELOAD
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c3_bench.json 2>gpurun_out/c3_bench.err
( OSVOS_FUSE_STAGE1=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip dp,gpu_reference,cpu_baseline,e2e_extra ) > gpurun_out/c3_bench_fuse_stage1.json 2>gpurun_out/c3_bench_fuse.err
( timeout 200 python scripts/time_forward.py ) > gpurun_out/c3_time_forward.txt 2>&1
( OSVOS_FUSE_STAGE1=1 timeout 200 python scripts/time_forward.py ) > gpurun_out/c3_time_forward_fuse.txt 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c3_launches_infer480.csv python scripts/one_forward.py ) > gpurun_out/c3_ncu.log 2>&1
( OSVOS_FUSE_STAGE1=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c3_launches_infer480_fuse.csv python scripts/one_forward.py ) > gpurun_out/c3_ncu_fuse.log 2>&1
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|sgd|cbce|sum_f32" -c 120 --csv --log-file gpurun_out/c3_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/c3_ncu_train.log 2>&1
for f in c3_new_kernels c3_pytest; do echo "== $f"; tail -5 gpurun_out/$f.txt; done
grep "^gated" gpurun_out/c3_gates.txt | cut -c1-300
cat gpurun_out/c3_ab_matrix.txt; tail -c 400 gpurun_out/c3_bench.err
