#!/bin/bash
# side_folded_wgrad v3b (TMA-fed ring, asynchronous window copies): parity + fwd+bwd bench + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_side_folded.py tests/test_gpu_objective.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/h_pytest.txt
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/h_train480.json 2> gpurun_out/h_train480.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|stage1|fold" -c 120 --csv --log-file gpurun_out/h_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/h_ncu_train.log 2>&1
( timeout 300 python bench.py --steps 20 --warmup 5 --skip cpu_baseline,gpu_reference,e2e_extra,roofline,parity ) > gpurun_out/h_bench_dp.json 2> gpurun_out/h_bench_dp.err
tail -4 gpurun_out/h_pytest.txt
head -c 330 gpurun_out/h_train480.json | tail -c 200; echo; tail -2 gpurun_out/h_train480.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/h_bench_dp.json"))
    print("value", d["value"], "dp fps", d["dp"]["fps"], "ms", d["dp"]["ms_per_step"], "parity", json.dumps(d["dp"].get("parity"))[:300])
except Exception as e:
    print("bench parse failed", e)
PY
grep "side_folded_wgrad" gpurun_out/h_launches_train480.csv | grep gpu__time | cut -d, -f1,5,12- | head
