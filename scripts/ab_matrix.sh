#!/bin/bash
# A/B of the library's opt-in switches on 480x854 inference, one process per switch (per-launch environment reads).
# Usage on a GPU box:  bash scripts/ab_matrix.sh > gpurun_out/ab_matrix.txt 2>&1   (about 15 s per line pair)
cd "$(dirname "$0")/.."
for sw in OSVOS_HALO_LEAN OSVOS_HALO_ST256 OSVOS_HALO_TMA_STORE OSVOS_SPLITK; do
  echo "== $sw (0 = default)"
  timeout 200 python scripts/ab_env.py $sw 0 1 --train || echo "FAILED: $sw"
done
echo "== OSVOS_HALO_LEAN 1 vs 2 (channel-split max pool)"
timeout 200 python scripts/ab_env.py OSVOS_HALO_LEAN 1 2 --train || echo "FAILED"
echo "== OSVOS_SPLITACC128 (1 = default)"
timeout 120 python scripts/ab_env.py OSVOS_SPLITACC128 1 0 || echo "FAILED"
