#!/bin/bash
mkdir -p gpurun_out
timeout 280 python tools/ab_kernels.py --pages 64 --steps 4 '{"B200MRC_IIRW_MODE":"trio"}' '{"B200MRC_IIRW_MODE":"trio","B200MRC_IIRW_DBG":"1"}' '{"B200MRC_IIRW_MODE":"trio","B200MRC_IIRW_DBG":"2"}' \
   '{"B200MRC_IIRW_MODE":"trio","B200MRC_IIRW_DBG":"3"}' '{"B200MRC_IIRW_MODE":"trio","B200MRC_IIRW_DBG":"7"}' > gpurun_out/q_ab4.log 2>&1; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/q_ab4.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['pages'], d['k_opt_iir_w'], d['k_opt_fir_w'], d['k_sauvola_mask'], d['total'], d['same_as_first'], d['env'])
PY
