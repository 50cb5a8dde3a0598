#!/bin/bash
# 240p fwd+bwd: per-op events (eager) and graphed step under both wgrad split rules
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python scripts/time_train.py 240 427 ) > gpurun_out/o_time_train_240.txt 2>&1
( timeout 300 python scripts/ab_env.py OSVOS_WGRAD_SPLITS new legacy 240 427 --train ) > gpurun_out/o_ab_splits_240.txt 2>&1
( timeout 300 python scripts/ab_env.py OSVOS_WGRAD_ROWS 1 0 240 427 --train ) > gpurun_out/o_ab_rows_240.txt 2>&1
grep -v "^  conv3x3 \|^  conv3x3_wgrad " gpurun_out/o_time_train_240.txt | tail -40
grep "fwd+bwd" gpurun_out/o_ab_splits_240.txt gpurun_out/o_ab_rows_240.txt
