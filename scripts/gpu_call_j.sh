#!/bin/bash
# multi-scale side wgrad launch: parity + benches (fwd+bwd, inference + dp leg) + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_side_folded.py tests/test_gpu_backward.py tests/test_gpu_objective.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/j_pytest.txt
( timeout 300 python bench.py --steps 20 --warmup 5 --workload train480 --skip cpu_baseline ) > gpurun_out/j_train480.json 2> gpurun_out/j_train480.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
( timeout 400 ncu --metrics $M --clock-control none -k regex:"conv|side|tail|wgrad|unpool|stage1|fold|sgd" -c 120 --csv --log-file gpurun_out/j_launches_train480.csv python scripts/one_train_step.py ) > gpurun_out/j_ncu_train.log 2>&1
( timeout 400 python bench.py --steps 20 --warmup 5 --skip cpu_baseline ) > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
tail -3 gpurun_out/j_pytest.txt
head -c 330 gpurun_out/j_train480.json | tail -c 200; echo; tail -2 gpurun_out/j_train480.err
grep "side_folded_wgrad\|side_grads_finish" gpurun_out/j_launches_train480.csv | grep gpu__time | cut -d, -f12- | head
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/j_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "launches", d["gpu_launches"], "dp fps", d["dp"]["fps"], "ms", d["dp"]["ms_per_step"])
    print(json.dumps(d["gpu_reference"])[:400])
except Exception as e:
    print("bench parse failed", e)
PY
