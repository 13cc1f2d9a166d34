#!/bin/bash
# Runs on the GPU box (via gpurun): round-2 measurement script.  Every step is bounded by its own timeout.
#   TAG=name          suffix of the output files        STEPS="smoke tests bench c2r1 launches full c3 c4 c5 osmo strong"
set -u
TAG=${TAG:-r02}
STEPS=${STEPS:-"smoke tests bench"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
nproc
has() { [[ " $STEPS " == *" $1 "* ]]; }
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
g=lambda k:(d.get(k) or {}).get('value')
d.setdefault('e2e_labeled', d.get('e2e') if 'labeled' in (d.get('e2e') or {}).get('api','') else None)
r=lambda v:None if v is None else round(v)
print('value',r(d['value']),'ms',round(d['ms_per_step'],3),'stages',{k:round(v,3) for k,v in d['stage_ms'].items()},'e2e',r(g('e2e')),'area_only',r(g('value_area_only')),r(g('e2e_area_only')),
      'auto',r(g('e2e_auto')),'png',r(g('e2e_png')),'auto_png',r(g('e2e_auto_png')),'cpu',d.get('cpu_baseline'),'diff',d.get('max_abs_diff_rgb_vs_cpu'),d.get('max_abs_diff_rgb_vs_cpu_labeled'))
print(' sustained',d.get('sustained'),'clocks',d.get('clocks'))
print(' e2e',d.get('e2e'))
print(' latency',d.get('latency_ms'),'roofline frac',d['roofline']['frac'],'local',d['roofline']['kernel_local_frac'])
PY
}
if has smoke; then echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; fi
if has tests; then echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8; fi
if has bench; then echo "== bench N=1 C2"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ${BENCH_FLAGS:-} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -2 gpurun_out/bench_${TAG}.err; summ gpurun_out/bench_${TAG}.json; fi
if has ref; then echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; tail -1 gpurun_out/bench_ref_${TAG}.err; cut -c1-600 gpurun_out/bench_ref_${TAG}.json; fi
for v in ${VARIANTS:-}; do
  echo "== variant $v"; timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline ${VARIANT_FLAGS:---skip-auto --skip-labeled} --min-seconds 0.5 --lib tools/dev/variants/$v.so > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; tail -1 gpurun_out/bench_$v.err
  summ gpurun_out/bench_$v.json
done
for w in c2r1 c3 c4 c5; do
  if has $w; then W=$(echo $w | sed 's/c2r1/C2r1/;s/c3/C3/;s/c4/C4/;s/c5/C5/'); echo "== bench $W"; timeout 900 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${w}_${TAG}.json 2> gpurun_out/bench_${w}_${TAG}.err; tail -1 gpurun_out/bench_${w}_${TAG}.err; summ gpurun_out/bench_${w}_${TAG}.json; fi
done
if has osmo; then echo "== bench C2 osmosnimki"; timeout 900 python bench.py --style osmosnimki --steps 10 --warmup 3 > gpurun_out/bench_osmo_${TAG}.json 2> gpurun_out/bench_osmo_${TAG}.err; tail -1 gpurun_out/bench_osmo_${TAG}.err; summ gpurun_out/bench_osmo_${TAG}.json; fi
if has launches; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --min-seconds 0 > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launch_${TAG}.log | cut -c1-300
fi
if has full; then
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${FULL_KERNELS:-raster_kernel|line_cover_kernel}" -s ${FULL_SKIP:-4} -c ${FULL_COUNT:-2} -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-auto --skip-labeled --min-seconds 0 > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
fi
if has labels_ncu; then
echo "== ncu launch list of the label kernels"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:label_ -c ${LABELS_NCU_COUNT:-90} --csv --log-file gpurun_out/launches_labels_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-auto --min-seconds 0 > gpurun_out/ncu_labels_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_labels_${TAG}.log | cut -c1-200
fi
if has full_labels; then
echo "== ncu full (label kernels)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${FULL_LABEL_KERNELS:-label_cover_kernel|label_curve_count_kernel|label_curve_expand_kernel}" -s ${FULL_LABEL_SKIP:-9} -c 3 -f -o gpurun_out/prof_labels_${TAG} \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-auto --min-seconds 0 > gpurun_out/ncu_full_labels_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_labels_${TAG}.log | cut -c1-200
fi
if has strong; then
echo "== strong scaling (under torchrun only)"
fi
ls -la gpurun_out | tail -20
