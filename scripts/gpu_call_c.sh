#!/bin/bash
# folded side-branch backward: parity tests, then fwd+bwd / parent-step A/B against the literal route
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_side_folded.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/c_pytest_new.txt
( timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_objective.py tests/test_gpu_optim.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/c_pytest_bwd.txt
for mode in folded literal; do
  ( OSVOS_SIDE_BWD=$mode timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/c_train480_$mode.json 2> gpurun_out/c_train480_$mode.err
done
( timeout 300 python bench.py --steps 20 --warmup 5 --skip cpu_baseline,gpu_reference,e2e_extra,roofline ) > gpurun_out/c_bench_dp.json 2> gpurun_out/c_bench_dp.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|stage1|fold" -c 120 --csv --log-file gpurun_out/c_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/c_ncu_train.log 2>&1
tail -5 gpurun_out/c_pytest_new.txt; tail -5 gpurun_out/c_pytest_bwd.txt
for mode in folded literal; do head -c 400 gpurun_out/c_train480_$mode.json; echo; tail -2 gpurun_out/c_train480_$mode.err; done
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c_bench_dp.json"))
    print("value", d["value"], "dp", json.dumps(d.get("dp"))[:600])
except Exception as e:
    print("bench parse failed", e)
PY
