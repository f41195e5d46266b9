#!/bin/bash
# Timing experiments on the tensor-core ALS kernel (FMS_TC_DBG bits: 1 one MMA per stage, 2 no probe-ahead, 4 no gathers,
# 8 no MMA, 16 no solve).  Results are WRONG by construction; only ms_users / ms_items matter.
for d in "$@"; do
  FMS_TC_DBG=$d timeout 200 python bench.py --skip-svm --skip-cpu --skip-e2e --skip-parity --steps 5 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg', $d, d['roofline']['iteration'])"
done
