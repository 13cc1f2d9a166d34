#!/bin/bash
# experiment (GPU box): labelled draw with the label stream at normal / high priority
for lp in ${LPS:-0 1}; do
  echo "== label_priority=$lp"
  timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-auto --min-seconds 0.5 --debug label_priority=$lp > gpurun_out/bench_lp$lp.json 2> gpurun_out/bench_lp$lp.err
  tail -1 gpurun_out/bench_lp$lp.err | cut -c1-200
  python - gpurun_out/bench_lp$lp.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
e=d['e2e']
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(e['value']),'ms',round(e['ms_per_step'],2),'label_device',round(e.get('ms_label_device',0),2),'stages',{k[:12]:round(v,2) for k,v in d['stage_ms'].items()})
PY
done
