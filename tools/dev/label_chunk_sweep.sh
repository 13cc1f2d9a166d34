#!/bin/bash
# experiment (GPU box): e2e of the labelled draw against the number of label chunks
for lc in ${LCS:-1 2 3}; do
  echo "== label_chunks=$lc"
  timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-auto --min-seconds 0.5 --debug label_chunks=$lc > gpurun_out/bench_lc$lc.json 2> gpurun_out/bench_lc$lc.err
  python - gpurun_out/bench_lc$lc.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
e=d['e2e']
print('value',round(d['value']),'e2e',round(e['value']),'ms',round(e['ms_per_step'],2),'label_device',round(e.get('ms_label_device',0),2),'area_only e2e',round((d.get('e2e_area_only') or {}).get('value',0)))
PY
done
