#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tee gpurun_out/pytest_t.log | tail -4
echo "=== report"; timeout 1500 python tools/report.py gpurun_out/report_t.json 9 > gpurun_out/report_t.log 2>&1; echo "report rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/report_t.json'))
for n,v in d['ct_external'].items(): print(n,{k:x.get('ms') for k,x in v.items()})
for n,v in d['r2c_c2r'].items(): print('real',n,{k:x.get('ms') for k,x in v.items() if 'ours' in k})
for n,v in d['ct_multiple'].items(): print('mult',n,{k:x['ms'] for k,x in v.items()})
PY
