#!/bin/bash
# users / items half-step time (+ the bench parity leg) per active-solver count:
#   tools/tc_cfg_sweep.sh 0:0 1:0 1:17 ...   (x : FMS_TC_NACT; 0 = default)
for spec in "$@"; do
  nact=${spec##*:}
  FMS_TC_NACT=$nact timeout 200 python bench.py --skip-svm --skip-cpu --skip-e2e --parity-rows 256 --steps 5 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg $cfg nact $nact', d['roofline']['iteration'], 'parity', d['parity']['max_rel'] if d.get('parity') else None)"
done
