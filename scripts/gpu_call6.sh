#!/bin/bash
# round-2 GPU call 6 (1 GPU): PDL for the inference pass, suspend-hint mbarrier waits; full suite, A/B, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -60 ) > gpurun_out/c6_pytest.txt
export OSVOS_ENV_RELOAD=1
( echo "== OSVOS_PDL_INFER (1 = default)"; timeout 200 python scripts/ab_env.py OSVOS_PDL_INFER 1 0 || echo FAILED
  echo "== OSVOS_WAIT_HINT_NS 0 vs 1000"; timeout 300 python scripts/ab_env.py OSVOS_WAIT_HINT_NS 0 1000 --train || echo FAILED
  echo "== OSVOS_WAIT_HINT_NS 0 vs 100"; timeout 200 python scripts/ab_env.py OSVOS_WAIT_HINT_NS 0 100 || echo FAILED
  echo "== OSVOS_WAIT_HINT_NS 0 vs 20000"; timeout 200 python scripts/ab_env.py OSVOS_WAIT_HINT_NS 0 20000 || echo FAILED
  for hw in "240 427" "720 1280"; do echo "== OSVOS_PDL_INFER at $hw"; timeout 200 python scripts/ab_env.py OSVOS_PDL_INFER 1 0 $hw || echo FAILED; echo "== OSVOS_WAIT_HINT_NS 0 vs 1000 at $hw"; timeout 200 python scripts/ab_env.py OSVOS_WAIT_HINT_NS 0 1000 $hw || echo FAILED; done
) > gpurun_out/c6_ab_matrix.txt 2>&1
unset OSVOS_ENV_RELOAD
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c6_bench.json 2>gpurun_out/c6_bench.err
( OSVOS_WAIT_HINT_NS=1000 timeout 300 python bench.py --steps 20 --warmup 5 --skip dp,gpu_reference,cpu_baseline,e2e_extra,parity ) > gpurun_out/c6_bench_hint1000.json 2>/dev/null
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( OSVOS_WAIT_HINT_NS=1000 timeout 300 ncu --metrics $M,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"conv|side|tail|stage1" -c 40 --csv --log-file gpurun_out/c6_launches_infer480_hint1000.csv python scripts/one_forward.py ) > gpurun_out/c6_ncu.log 2>&1
tail -4 gpurun_out/c6_pytest.txt; cat gpurun_out/c6_ab_matrix.txt; tail -c 300 gpurun_out/c6_bench.err; head -c 600 gpurun_out/c6_bench.json
