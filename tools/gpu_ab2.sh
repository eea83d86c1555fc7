#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ab_kernels.py --pages 56 --steps 4 '{"B200MRC_IIRW_MODE":"single"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_LAG":"0"}' '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_LAG":"1"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_LAG":"2"}' '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_LAG":"3"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_LAG":"4"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_STAGES":"10","B200MRC_IIRW_LAG":"2"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_STAGES":"10","B200MRC_IIRW_LAG":"4"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_STAGES":"10","B200MRC_IIRW_LAG":"6"}' \
  '{"B200MRC_IIRW_MODE":"pair","B200MRC_IIRW_STAGES":"10","B200MRC_IIRW_LAG":"8"}' > gpurun_out/q2_ab.log 2>&1; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/q2_ab.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['k_opt_iir_w'], d['total'], d['same_as_first'], d['env'])
PY
B200MRC_IIRW_MODE=pair timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_opt_iir_w2 -c 1 -o gpurun_out/r1q_pair -f python tools/profile_step.py --pages 64 --warmup 0 --steps 1 > gpurun_out/ncu_q.log 2>&1; tail -3 gpurun_out/ncu_q.log
